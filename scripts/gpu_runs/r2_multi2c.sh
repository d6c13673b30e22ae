# round 2: two GPUs, units off a counter + NCCL on a high-priority stream: is the gather still at the mercy of who gets the SMs first?
mkdir -p gpurun_out
i=0
for sp in 0 0 0 8 8; do
i=$((i+1))
B200Z_FAST_SPARE_SMS=$sp timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_n2c_$i.json 2> gpurun_out/bench_n2c_$i.err
python - <<PY
import json
for ln in open('gpurun_out/bench_n2c_$i.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print("spare $sp: value", round(d['value'],1), "ms", round(d['ms_per_step'],2), "decode_only", round(d['decode_only']['ms_per_step'],2), "strong", round(d['strong']['value'],1), round(d['strong']['ms_per_step'],2))
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1c.json')); print('N1: value', d['value'], d['ms_per_step'])"
