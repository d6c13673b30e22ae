# round 2, call 19: walk_order in shared memory, rle_emit with 8-byte stores; deflate fast batch timings; launch list of config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -4
timeout 900 python scripts/bench_defl_fast.py 256 4 1 2>&1 | tail -3
timeout 900 python scripts/bench_defl_fast.py 1024 4 1 2>&1 | tail -3
timeout 600 python scripts/bench_defl_fast.py 64 4 3 2>&1 | tail -3
timeout 900 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/bench_c4_r2_19.json 2> gpurun_out/bench_c4_r2_19.err; tail -2 gpurun_out/bench_c4_r2_19.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_19.json')); print('config 4: value', d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2_launches_r2_19.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bz2_ncu19.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2_launches_r2_19.csv 14 2>&1 | tail -18
