# round 2, call 33: k_defl_match with 4096 positions per CTA (two CTAs per SM)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_defl6.py 64 6 2>&1 | tail -1
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_r2_33.json 2> gpurun_out/bench_c3_r2_33.err; tail -1 gpurun_out/bench_c3_r2_33.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_r2_33.json')); print('config 3: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
