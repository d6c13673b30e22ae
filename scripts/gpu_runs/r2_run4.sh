# round 2, call 4: k_inflate_fast with per-byte finality tracking in the LZ77 pass (all threads)
mkdir -p gpurun_out
export B200Z_FAST=1
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_fast3.json 2> gpurun_out/bench_fast3.err; tail -c 900 gpurun_out/bench_fast3.json; tail -3 gpurun_out/bench_fast3.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_fast -s 3 -c 1 -o gpurun_out/r2_fast_v3 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/ncu_fast_v3.log 2>&1
tail -2 gpurun_out/ncu_fast_v3.log
