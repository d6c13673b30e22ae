# round 2: two GPUs, final tree -- multi-GPU tests, the default line under torchrun, config 4 sharded
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | grep -v Warn | tail -2 | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err; tail -1 gpurun_out/bench_n2_final.err | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/bench_n2_final.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print("N2 value", d['value'], "ms", d['ms_per_step'], "decode_only", d['decode_only']['value'], "strong", d['strong']['value'], "e2e", d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config 4 --no-cpu-baseline > gpurun_out/bench_c4_n2_final.json 2> gpurun_out/bench_c4_n2_final.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_c4_n2_final.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print("config 4 N2:", d['value'], d['unit'], d['ms_per_step'], 'ms')
PY
