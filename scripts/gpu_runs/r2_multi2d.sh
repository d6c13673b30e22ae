# round 2 (re-entry): two GPUs on the final tree -- multi-GPU tests, bench.py under torchrun (both arms), config 4 sharded
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi2.log
grep -v Warn gpurun_out/pytest_multi2.log | tail -3 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-side-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print("N2 value", d['value'], "ms", d['ms_per_step'], "decode_only", d['decode_only'], "strong", d['strong']['value'], d['strong']['ms_per_step'], "e2e", d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    for r in d['per_rank']: print(r)
except Exception as e: print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4 --no-cpu-baseline > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; cut -c1-300 gpurun_out/bench_c4_n2.json; tail -2 gpurun_out/bench_c4_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-side-configs > gpurun_out/bench_n1_pair.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_pair.json')); print('N1 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])"
