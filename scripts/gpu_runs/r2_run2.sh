# round 2, call 2: why is k_inflate_fast slow -- ncu with source counters
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_inflate_fast -s 3 -c 1 -o gpurun_out/r2_fast_v1 -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fast_v1.log 2>&1
tail -5 gpurun_out/ncu_fast_v1.log
ls -la gpurun_out/*.ncu-rep
