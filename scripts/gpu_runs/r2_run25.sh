# round 2, call 25: K8 walks as one pass into slots + a coalesced placement pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py tests/test_zip_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -4
timeout 900 python bench.py --config 4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_r2_25.json 2> gpurun_out/bench_c4_r2_25.err; tail -1 gpurun_out/bench_c4_r2_25.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_25.json')); print('config 4: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2_launches_r2_25.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bz2_ncu24.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2_launches_r2_25.csv 40 2>&1 | grep -v "bz2e::" | head -24
