# round 2, call 28: k_defl_match with the four-byte quick reject; zip extraction in chunks (config 5)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_deflate_gpu.py tests/test_zip_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_defl6.py 64 6 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/defl6_launches_r2_28.csv python scripts/bench_defl6.py 64 6 > gpurun_out/defl6_ncu28.log 2>&1
python scripts/launch_summary.py gpurun_out/defl6_launches_r2_28.csv 3 2>&1 | tail -3
timeout 900 python bench.py --config 3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_r2_28.json 2> gpurun_out/bench_c3_r2_28.err; tail -1 gpurun_out/bench_c3_r2_28.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c3_r2_28.json')); print('config 3: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
for ch in 1 8; do
B200Z_ZIP_CHUNKS=$ch timeout 900 python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_r2_28_$ch.json 2> gpurun_out/bench_c5_r2_28_$ch.err; tail -1 gpurun_out/bench_c5_r2_28_$ch.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c5_r2_28_$ch.json')); print('config 5 chunks $ch: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
done
