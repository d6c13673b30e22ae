mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_inflate_gpu.py -x -q > gpurun_out/pytest_infl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_infl.log
grep -v Warn gpurun_out/pytest_infl.log | tail -4 | cut -c1-250
for cfg in "8 0" "4 0" "16 0" "8 1"; do
  set -- $cfg
  B200Z_UPW=$1 B200Z_SPEC_G=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s_$1_$2.json 2> gpurun_out/bench_s_$1_$2.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_s_$1_$2.json').read().strip().splitlines()[-1])
print('upw=$1 G=$2', round(d['value'],1), d['roofline']['kernels'])
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate_decode -s 3 -c 1 -o gpurun_out/r1_decode_spec2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_decode_spec2.log 2>&1
ls -la gpurun_out/r1_decode_spec2.ncu-rep
