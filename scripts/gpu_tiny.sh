timeout 300 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py -x -q 2>&1 | tail -1 | cut -c1-120
