mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
grep -v Warn gpurun_out/pytest_gpu_all.log | tail -12 | cut -c1-250
