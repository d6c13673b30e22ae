# round 2: two GPUs, the new default kernel -- how many SMs to leave to the all-gather
mkdir -p gpurun_out
for sp in 0 4 8 16 24; do
B200Z_FAST_SPARE_SMS=$sp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((sp % 10)) bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n2_sp$sp.json 2> gpurun_out/bench_n2_sp$sp.err
python - <<PY
import json
for ln in open('gpurun_out/bench_n2_sp$sp.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print("spare $sp: value", round(d['value'],1), "ms", round(d['ms_per_step'],2), "decode_only", round(d['decode_only']['ms_per_step'],2), "strong", round(d['strong']['value'],1), round(d['strong']['ms_per_step'],2))
PY
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_full.json 2> gpurun_out/bench_n2_full.err
python - <<PY
import json
for ln in open('gpurun_out/bench_n2_full.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print("N2 full: value", d['value'], d['ms_per_step'], "e2e", d['e2e']['value'], d['e2e']['ms_per_step'], "cpu", d['cpu_baseline']['value'])
        for r in d['per_rank']: print(r)
PY
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi2b.log 2>&1; grep -v Warn gpurun_out/pytest_multi2b.log | tail -2
