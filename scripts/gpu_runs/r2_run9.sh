# round 2, call 9: the new defaults on one device -- full bench line, the pair on another rank's data, launch list
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_9.json 2> gpurun_out/bench_9.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_9.json'))
print("value",d['value'],d['ms_per_step'],"e2e",d['e2e']['value'], d['e2e']['ms_per_step'],"alt",d['alt_kernel'])
print("roofline", d['roofline']['frac'], d['roofline']['kernels'], "cpu", d['cpu_baseline']['value'])
for k,v in d['configs'].items():
    if 'error' in v: print(k, v); continue
    print(k, "value", round(v['value'],2), "ms", round(v['ms_per_step'],1), "cpu", round(v['cpu_baseline']['value'],3), "parity", v['parity'], "wall", v.get('wall_s'))
PY
tail -3 gpurun_out/bench_9.err
B200Z_BENCH_STREAM0=4096 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_9_s4096.json 2> gpurun_out/bench_9_s4096.err; python -c "
import json; d=json.load(open('gpurun_out/bench_9_s4096.json')); print('stream 4096: value', d['value'], d['ms_per_step'], 'alt', d['alt_kernel']['value'], d['alt_kernel']['ms_per_step'])"
B200Z_GZIP_PIPED_WALK=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-side-configs > gpurun_out/bench_9_nopipe.json 2> gpurun_out/bench_9_nopipe.err; python -c "
import json; d=json.load(open('gpurun_out/bench_9_nopipe.json')); print('walk first: e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-configs > gpurun_out/bench_9_ncu.log 2>&1; tail -1 gpurun_out/bench_9_ncu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke9.log 2>&1; tail -2 gpurun_out/smoke9.log
