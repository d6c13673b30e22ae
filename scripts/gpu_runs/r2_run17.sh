# round 2, call 17: ncu --set full of k_bz2_entropy_fast on a single 900 kB block (one CTA): where do the clocks of a group go?
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bz2_entropy_fast -c 1 -o gpurun_out/bz2fast_r2_17 -f python scripts/bench_bz2_small.py 900000 > gpurun_out/bz2fast_ncu.log 2>&1
tail -3 gpurun_out/bz2fast_ncu.log
ls -la gpurun_out/
