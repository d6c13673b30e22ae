# round 2, call 34: k_defl_match with 1024 threads per CTA (64 warps per SM)
mkdir -p gpurun_out
timeout 300 python scripts/bench_defl6.py 64 6 2>&1 | tail -1
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_r2_34.json 2> gpurun_out/bench_c3_r2_34.err; tail -1 gpurun_out/bench_c3_r2_34.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_r2_34.json')); print('config 3: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
