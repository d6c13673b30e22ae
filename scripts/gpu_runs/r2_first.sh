# First GPU call of the next round: everything written after round 1's GPU budget ran out, in order of risk.
# (gpurun --timeout 2400 -- 'bash scripts/gpu_runs/r2_first.sh')
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_all.log
grep -v Warn gpurun_out/pytest_gpu_all.log | tail -12 | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
timeout 900 python scripts/bench_file.py > gpurun_out/bench_file.json 2> gpurun_out/bench_file.err; tail -c 2500 gpurun_out/bench_file.json; tail -3 gpurun_out/bench_file.err
timeout 600 python scripts/parity_survey.py > gpurun_out/parity_survey.jsonl 2> gpurun_out/parity_survey.err; cat gpurun_out/parity_survey.jsonl | cut -c1-200
