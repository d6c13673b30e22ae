# round 2, call 15 (re-entry): state of the tree -- GPU tests, smoke, default bench line (side configs 3/4/5), single-block bzip2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 ) 2>&1 | tail -9
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_r2_15.json 2> gpurun_out/bench_r2_15.err; tail -3 gpurun_out/bench_r2_15.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_15.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d.get('e2e'),'roof',d.get('roofline'))
print('cpu',d.get('cpu_baseline'))
for k,v in (d.get('configs') or {}).items():
    print(k, json.dumps(v)[:700])
PY
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -5
