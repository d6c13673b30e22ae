mkdir -p gpurun_out
for cfg in "8 0" "8 1" "4 0"; do
  set -- $cfg
  B200Z_LIB=$PWD/archive_b200/libb200z_nostore.so B200Z_UPW=$1 B200Z_SPEC_G=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_ns_$1_$2.json 2> gpurun_out/bench_ns_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_ns_$1_$2.json').read().strip().splitlines()[-1])
    print('NOSTORE upw=$1 G=$2', round(d['value'],1), d['roofline']['kernels'])
except Exception as e:
    print('fail', e); print(open('gpurun_out/bench_ns_$1_$2.err').read()[-600:])
PY
done
