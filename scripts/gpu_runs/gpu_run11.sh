mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py -x -q > gpurun_out/pytest_infl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_infl.log
for u in 16 8 4 2; do B200Z_UPW=$u timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_upw$u.log 2>&1; done
grep -v Warn gpurun_out/pytest_infl.log | tail -3
