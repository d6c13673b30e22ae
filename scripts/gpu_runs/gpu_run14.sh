mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_deflate_gpu.py -x -q > gpurun_out/pytest_defl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_defl.log
grep -v Warn gpurun_out/pytest_defl.log | tail -12 | cut -c1-220
