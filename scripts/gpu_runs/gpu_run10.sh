mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py -x -q > gpurun_out/pytest_bz2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bz2.log
DEFL_MIB=1 DEFL_CHECK_ORACLE=0 timeout 900 python scripts/bench_codecs.py > gpurun_out/bench_codecs_bz.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 6 -c 2 -f -o gpurun_out/prof_inflate python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
grep -v Warn gpurun_out/pytest_bz2.log | tail -4 | cut -c1-200; tail -2 gpurun_out/bench_codecs_bz.log | cut -c1-600
