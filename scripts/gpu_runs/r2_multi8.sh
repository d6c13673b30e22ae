# round 2 (re-entry): eight GPUs on the final tree -- bench.py exactly as the driver launches it (both arms)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -2 gpurun_out/bench_n8.err | cut -c1-300
python - <<'PY'
import json
for ln in open('gpurun_out/bench_n8.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        print("N8 value", d['value'], "ms", d['ms_per_step'], "decode_only", d['decode_only'], "strong", d['strong']['value'], d['strong']['ms_per_step'], "e2e", d['e2e'], 'cpu', d.get('cpu_baseline',{}).get('value'))
        for r in d['per_rank']: print(r['rank'], round(r['ms_per_step'],2), round(r['decode_only_ms_per_step'],2), round(r['k_inflate_fast_ms'],2), r['sm_mhz'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --impl reference > gpurun_out/bench_n8_ref.json 2> gpurun_out/bench_n8_ref.err; cut -c1-400 gpurun_out/bench_n8_ref.json
