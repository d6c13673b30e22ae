# round 2, call 27: k_defl_match as a per-lane state machine (positions off a counter, four-byte compares): parity, level 6 timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_zz_deflate_level0_windows_gpu.py tests/test_bzip2_enc_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_defl6.py 64 6 2>&1 | tail -1
timeout 300 python scripts/bench_defl6.py 64 9 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/defl6_launches_r2_27.csv python scripts/bench_defl6.py 64 6 > gpurun_out/defl6_ncu27.log 2>&1
python scripts/launch_summary.py gpurun_out/defl6_launches_r2_27.csv 8 2>&1 | tail -10
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_r2_27.json 2> gpurun_out/bench_c3_r2_27.err; tail -1 gpurun_out/bench_c3_r2_27.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_r2_27.json')); print('config 3: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
