mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_bzip2_gpu.py -x -q -k "fixtures or cat_jpg" > gpurun_out/sanitizer_bz2.log 2>&1; echo "san rc=$?" >> gpurun_out/sanitizer_bz2.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
for u in 32 16 8 4; do B200Z_UPW=$u timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_upw$u.log 2>&1; done
for b in 2 4 8; do B200Z_EXPAND_BPS=$b timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_bps$b.log 2>&1; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 6 -c 2 -f -o gpurun_out/prof_inflate python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -15 gpurun_out/pytest.log; tail -2 gpurun_out/bench.log | cut -c1-1200
