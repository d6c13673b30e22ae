# round 2, call 26: Deflate level 6 -- launch list (64 MiB) and ncu --set full of k_defl_match
mkdir -p gpurun_out
timeout 300 python scripts/bench_defl6.py 64 6 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/defl6_launches_r2_26.csv python scripts/bench_defl6.py 64 6 > gpurun_out/defl6_ncu26.log 2>&1
python scripts/launch_summary.py gpurun_out/defl6_launches_r2_26.csv 20 2>&1 | tail -22
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_defl_match -c 1 -o gpurun_out/defl_match_r2_26 -f python scripts/bench_defl6.py 64 6 > gpurun_out/defl_match_ncu26.log 2>&1
tail -2 gpurun_out/defl_match_ncu26.log
