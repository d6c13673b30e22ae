mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
grep -v Warn gpurun_out/pytest_gpu_all.log | tail -3 | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/bench_launches.csv 6 | tee gpurun_out/bench_launch_summary.md
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/r1_inflate_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_inflate_final.log 2>&1
ls -la gpurun_out/r1_inflate_final.ncu-rep
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
