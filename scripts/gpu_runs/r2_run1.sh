# round 2, call 1: the shared-memory inflate kernel's first contact with the hardware
mkdir -p gpurun_out
(which dart flutter; ls /usr/lib/dart /opt/dart-sdk 2>&1 | head -3) > gpurun_out/dart_probe.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> gpurun_out/dart_probe.txt
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py tests/test_zz_gzip_stream_semantics_gpu.py -x -q -m gpu > gpurun_out/pytest_inflate.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_inflate.log
grep -v Warn gpurun_out/pytest_inflate.log | tail -8 | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_fast.json 2> gpurun_out/bench_fast.err; tail -c 1800 gpurun_out/bench_fast.json; tail -5 gpurun_out/bench_fast.err
B200Z_FAST=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err; tail -c 600 gpurun_out/bench_old.json
