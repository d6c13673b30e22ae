mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py -x -q > gpurun_out/pytest_i.log 2>&1; tail -1 gpurun_out/pytest_i.log | cut -c1-120
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cs.json 2> gpurun_out/bench_cs.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cs.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['kernels'])
PY
