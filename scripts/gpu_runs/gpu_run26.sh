mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zip_gpu.py -x -q > gpurun_out/pytest_zip.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_zip.log
grep -v Warn gpurun_out/pytest_zip.log | tail -12 | cut -c1-250
ZIP_MEMBERS=256 ZIP_FLUSH=1 timeout 900 python scripts/bench_zip.py 2>&1 | tail -2 | cut -c1-600 | tee gpurun_out/bench_zip_flush.json
ZIP_MEMBERS=256 ZIP_FLUSH=0 timeout 900 python scripts/bench_zip.py 2>&1 | tail -2 | cut -c1-600 | tee gpurun_out/bench_zip_noflush.json
