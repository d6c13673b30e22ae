mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_inflate_gpu.py -x -q > gpurun_out/pytest_infl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_infl.log
grep -v Warn gpurun_out/pytest_infl.log | tail -8 | cut -c1-250
for g in 0 1 2; do
  B200Z_SPEC_G=$g timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_spec_g$g.json 2> gpurun_out/bench_spec_g$g.err
  echo "SPEC_G=$g"; cut -c1-700 gpurun_out/bench_spec_g$g.json; tail -2 gpurun_out/bench_spec_g$g.err | cut -c1-300
done
B200Z_UPW=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spec_upw4.json 2>&1; echo UPW4; cut -c1-500 gpurun_out/bench_spec_upw4.json
B200Z_UPW=16 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_spec_upw16.json 2>&1; echo UPW16; cut -c1-500 gpurun_out/bench_spec_upw16.json
