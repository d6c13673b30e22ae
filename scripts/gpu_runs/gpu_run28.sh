mkdir -p gpurun_out
for bps in 2 3 4 6 8 16; do
  B200Z_EXPAND_BPS=$bps timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_bps_$bps.json 2> gpurun_out/bench_bps_$bps.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_bps_$bps.json').read().strip().splitlines()[-1])
print('bps=$bps', round(d['value'],1), d['roofline']['kernels'])
PY
done
timeout 1200 python -m pytest tests/test_zip_gpu.py -x -q 2>&1 | tail -2 | cut -c1-200
