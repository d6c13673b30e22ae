mkdir -p gpurun_out
timeout 300 python scripts/dbg_bz2.py > gpurun_out/dbg_bz2.log 2>&1
timeout 1200 python scripts/bench_codecs.py > gpurun_out/bench_codecs.log 2>&1; echo "rc=$?" >> gpurun_out/bench_codecs.log
tail -30 gpurun_out/dbg_bz2.log | cut -c1-220; tail -5 gpurun_out/bench_codecs.log | cut -c1-1500
