mkdir -p gpurun_out
for cfg in "8 0" "4 0" "16 0" "8 2"; do
  set -- $cfg
  B200Z_UPW=$1 B200Z_SPEC_G=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s_$1_$2.json 2> gpurun_out/bench_s_$1_$2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s_$1_$2.json').read().strip().splitlines()[-1])
print('upw=$1 G=$2', round(d['value'],1), d['roofline']['kernels'])
PY
done
