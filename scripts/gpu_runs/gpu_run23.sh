mkdir -p gpurun_out
N_UNITS=16384 timeout 600 python scripts/dbg_pieces.py 2>&1 | tail -9
timeout 1500 python -m pytest tests/test_inflate_gpu.py -x -q > gpurun_out/pytest_infl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_infl.log
grep -v Warn gpurun_out/pytest_infl.log | tail -3 | cut -c1-250
for cfg in "8 0" "4 0"; do
  set -- $cfg
  B200Z_UPW=$1 B200Z_SPEC_G=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s_$1_$2.json 2> gpurun_out/bench_s_$1_$2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s_$1_$2.json').read().strip().splitlines()[-1])
print('upw=$1 G=$2', round(d['value'],1), d['roofline']['kernels'], 'e2e', d['e2e']['value'])
PY
done
