mkdir -p gpurun_out; nvidia-smi > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
for u in 32 16 8 4; do B200Z_UPW=$u timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_upw$u.log 2>&1; done
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_inflate_gpu.py -x -q -k "mixed or bad_data" > gpurun_out/sanitizer.log 2>&1; echo "san rc=$?" >> gpurun_out/sanitizer.log
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench.log
