# round 2, call 16: k_bz2_entropy_fast -- parity on the GPU, single block / 64 MiB / config 4 timings, launch list of the decode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -6
B200Z_BZ2_FAST=0 timeout 300 python scripts/bench_bz2_small.py 900000 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2_launches_r2_16.csv python scripts/bench_bz2_small.py 67108864 > gpurun_out/bz2_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2_launches_r2_16.csv 2>&1 | tail -40
timeout 900 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/bench_c4_r2_16.json 2> gpurun_out/bench_c4_r2_16.err; tail -2 gpurun_out/bench_c4_r2_16.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_16.json')); print('config 4: value', d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
