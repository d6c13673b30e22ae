# round 2, call 7: the state of everything that changed, on one device (~12 min)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py tests/test_zz_gzip_stream_semantics_gpu.py tests/test_bzip2_gpu.py tests/test_multi_gpu.py tests/test_zz_bzip2_damaged_gpu.py -x -q -m gpu > gpurun_out/pytest_7.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_7.log
grep -v Warn gpurun_out/pytest_7.log | tail -4 | cut -c1-250
B200Z_FAST=1 timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_7_fast.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_7_fast.log
grep -v Warn gpurun_out/pytest_7_fast.log | tail -3 | cut -c1-250
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_7.json 2> gpurun_out/bench_7.err; cut -c1-1500 gpurun_out/bench_7.json; tail -3 gpurun_out/bench_7.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_bz2dec_launches.csv python scripts/bench_bz2_small.py > gpurun_out/bz2_small.log 2>&1; tail -2 gpurun_out/bz2_small.log
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_7_ref.json 2> gpurun_out/bench_7_ref.err; cut -c1-400 gpurun_out/bench_7_ref.json
