mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bzip2_enc_gpu.py -x -q 2>&1 | tail -2 | cut -c1-200
BZE_MIB=256 timeout 900 python scripts/bench_bz2enc.py > gpurun_out/bz2enc_bench.json 2> gpurun_out/bz2enc_bench.err; tail -n 3 gpurun_out/bz2enc_bench.json | cut -c1-400
BZE_MIB=64 BZE_CHECK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2enc_launches.csv python scripts/bench_bz2enc.py > gpurun_out/bz2enc_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2enc_launches.csv 12 | tee gpurun_out/bz2enc_launch_summary.md
