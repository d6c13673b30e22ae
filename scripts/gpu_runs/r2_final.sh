# round 2, final validation of the tree: all GPU tests, smoke(), both bench arms at N = 1 (default flags), launch list of the default step
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'roof',d['roofline']['frac'],'launches',d.get('gpu_launches'),'clocks',d.get('clocks'))
print('cpu',d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
for k,v in (d.get('configs') or {}).items():
    print(k, v.get('metric'), round(v.get('value',0),3), v.get('unit'), 'ms', round(v.get('ms_per_step',0),1), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), 'parity', v.get('parity') or v.get('check'))
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-300 gpurun_out/bench_final_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/b_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/bench_launches_r2_final.csv 8 2>&1 | tail -10
