mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py -x -q 2>&1 | tail -2 | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_exp.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['roofline']['kernels'], 'e2e', d['e2e']['value'])
PY
BZ_MIB=512 timeout 600 python scripts/bench_bz2_multi.py 2>&1 | tail -1 | tee gpurun_out/bz2_multi_1.json
