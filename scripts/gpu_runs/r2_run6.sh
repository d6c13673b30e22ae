# round 2, call 6: k_inflate_fast v5 (warp-converged pass loops, second-level codes inside the bulk loop) -- parity + timing + profile
mkdir -p gpurun_out
export B200Z_FAST=1
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zip_gpu.py tests/test_zz_gzip_stream_semantics_gpu.py -x -q -m gpu > gpurun_out/pytest_inflate6.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_inflate6.log
grep -v Warn gpurun_out/pytest_inflate6.log | tail -4 | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_fast6.json 2> gpurun_out/bench_fast6.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fast6.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"; tail -3 gpurun_out/bench_fast6.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_fast -s 3 -c 1 -o gpurun_out/r2_fast_v6 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/ncu_fast_v6.log 2>&1
tail -2 gpurun_out/ncu_fast_v6.log
