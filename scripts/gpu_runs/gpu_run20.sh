mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate_decode -s 3 -c 1 -o gpurun_out/r1_decode_spec -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_decode_spec.log 2>&1
tail -3 gpurun_out/ncu_decode_spec.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
