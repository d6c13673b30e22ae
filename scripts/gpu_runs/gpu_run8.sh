mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bzip2_gpu.py -x -q > gpurun_out/pytest_bz2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bz2.log
DEFL_MIB=64 BZ_MIB=0 DEFL_CHECK_ORACLE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_deflate.csv python scripts/bench_codecs.py > gpurun_out/ncu_defl.log 2>&1
DEFL_MIB=1 BZ_MIB=128 DEFL_CHECK_ORACLE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bzip2.csv python scripts/bench_codecs.py > gpurun_out/ncu_bz2.log 2>&1
grep -v Warn gpurun_out/pytest_bz2.log | tail -5
