# round 2, call 29: k_defl_fast_batch with 32 positions per turn: parity on the device, batch and single-stream timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_zz_file_codec_gpu.py tests/test_zip_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python scripts/bench_defl_fast.py 256 4 1 2>&1 | tail -2
timeout 900 python scripts/bench_defl_fast.py 1024 4 1 2>&1 | tail -2
timeout 600 python scripts/bench_defl_fast.py 64 4 3 2>&1 | tail -2
timeout 600 python scripts/bench_defl_fast.py 64 4 2 2>&1 | tail -2
