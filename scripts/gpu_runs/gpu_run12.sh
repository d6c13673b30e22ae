mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
for u in 16 8 4; do B200Z_UPW=$u timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_upw$u.log 2>&1; done
DEFL_MIB=1 DEFL_CHECK_ORACLE=0 timeout 900 python scripts/bench_codecs.py > gpurun_out/bench_codecs_bz.log 2>&1
grep -v Warn gpurun_out/pytest.log | tail -3; tail -2 gpurun_out/bench_codecs_bz.log | cut -c1-500
