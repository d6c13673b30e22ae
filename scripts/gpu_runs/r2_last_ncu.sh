# round 2, last call: ncu --set full of k_inflate_fast in the final build (one launch of bench.py's batch)
mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:k_inflate_fast -s 3 -c 1 -o gpurun_out/inflate_fast_final -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/inflate_fast_final_ncu.log 2>&1
tail -2 gpurun_out/inflate_fast_final_ncu.log | cut -c1-200
