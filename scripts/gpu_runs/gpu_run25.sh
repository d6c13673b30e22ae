mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zip_gpu.py -x -q > gpurun_out/pytest_zip.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_zip.log
grep -v Warn gpurun_out/pytest_zip.log | tail -25 | cut -c1-250
ZIP_MEMBERS=256 timeout 900 python scripts/bench_zip.py 2>&1 | tail -3 | cut -c1-600 | tee gpurun_out/bench_zip.json
