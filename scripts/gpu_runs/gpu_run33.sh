mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_bzip2_shard.py -x -q -m gpu 2>&1 | grep -v Warn | tail -8 | cut -c1-220
