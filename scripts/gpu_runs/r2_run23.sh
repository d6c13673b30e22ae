# round 2, call 23: k_bz2_entropy_fast with three warps (walker: start bits only / decoder + symbolic MTF / recorder)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -4
timeout 900 python bench.py --config 4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_r2_23.json 2> gpurun_out/bench_c4_r2_23.err; tail -1 gpurun_out/bench_c4_r2_23.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_23.json')); print('config 4: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bz2_entropy_fast -c 1 -o gpurun_out/bz2fast_r2_23 -f python scripts/bench_bz2_small.py 900000 > gpurun_out/bz2fast_ncu23.log 2>&1
tail -2 gpurun_out/bz2fast_ncu23.log
