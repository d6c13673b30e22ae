mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_bzip2_enc_gpu.py -x -q > gpurun_out/pytest_bz2e.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bz2e.log
grep -v Warn gpurun_out/pytest_bz2e.log | tail -15 | cut -c1-250
BZE_MIB=256 timeout 900 python scripts/bench_bz2enc.py > gpurun_out/bz2enc_bench.json 2> gpurun_out/bz2enc_bench.err; tail -3 gpurun_out/bz2enc_bench.json gpurun_out/bz2enc_bench.err | cut -c1-400
BZE_MIB=64 BZE_CHECK=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2enc_launches.csv python scripts/bench_bz2enc.py > gpurun_out/bz2enc_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2enc_launches.csv 30 | tee gpurun_out/bz2enc_launch_summary.md
