# round 2, very last call: the default bench line (no side configs) on the final tree -- clock samples every 25 ms
timeout 110 python bench.py --no-side-configs --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -1 gpurun_out/bench_last.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'], 'launches', d['gpu_launches'])"
