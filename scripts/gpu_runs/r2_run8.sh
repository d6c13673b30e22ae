# round 2, call 8: k_inflate_fast with the funnel-shift loops; why config 3 came out at 0.19 GB/s inside bench.py
mkdir -p gpurun_out
B200Z_FAST=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_fast8.json 2> gpurun_out/bench_fast8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_fast8.json')); print('FAST', d['value'], d['ms_per_step'], d['roofline']['kernels'])"; tail -2 gpurun_out/bench_fast8.err
B200Z_FAST=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inflate_fast -s 3 -c 1 -o gpurun_out/r2_fast_v8 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/ncu_fast_v8.log 2>&1
tail -1 gpurun_out/ncu_fast_v8.log
B200Z_FAST=1 timeout 600 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py -x -q -m gpu > gpurun_out/pytest_8_fast.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_8_fast.log; grep -v Warn gpurun_out/pytest_8_fast.log | tail -2
BZ_MIB=0 DEFL_CHECK_ORACLE=0 timeout 300 python scripts/bench_codecs.py > gpurun_out/c3_old_script.json 2>&1; tail -c 400 gpurun_out/c3_old_script.json
timeout 600 python bench.py --config 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-500 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_defl6_launches.csv python -c "
import os; os.environ['DEFL_MIB']='64'; os.environ['BZ_MIB']='0'; os.environ['DEFL_CHECK_ORACLE']='0'
exec(open('scripts/bench_codecs.py').read())" > gpurun_out/c3_ncu.log 2>&1; tail -1 gpurun_out/c3_ncu.log
