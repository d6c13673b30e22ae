# round 2, call 30: window turn with the four-byte quick reject in the per-lane chain walk
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_zz_file_codec_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python scripts/bench_defl_fast.py 256 4 1 2>&1 | tail -2
timeout 600 python scripts/bench_defl_fast.py 64 4 3 2>&1 | tail -2
timeout 600 python scripts/bench_defl_fast.py 64 4 2 2>&1 | tail -2
