mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_bzip2_gpu.py -x -q > gpurun_out/pytest_bz2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bz2.log
timeout 1500 python -m pytest tests/test_deflate_gpu.py -x -q > gpurun_out/pytest_defl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_defl.log
timeout 900 python -m pytest tests/test_inflate_gpu.py -x -q > gpurun_out/pytest_infl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_infl.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_deflate_gpu.py -x -q -k "level_0 or framing" > gpurun_out/sanitizer_defl.log 2>&1; echo "san rc=$?" >> gpurun_out/sanitizer_defl.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
for f in bz2 defl infl; do grep -v Warn gpurun_out/pytest_$f.log | tail -12 | cut -c1-250; done; tail -4 gpurun_out/sanitizer_defl.log; tail -2 gpurun_out/bench.log | cut -c1-900
