# round 2, call 22: K8 walk kernels with a work counter; config 4; launch list; default inflate build back to the per-byte LZ77 pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py tests/test_inflate_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -4
timeout 900 python bench.py --config 4 --steps 3 --warmup 1 > gpurun_out/bench_c4_r2_22.json 2> gpurun_out/bench_c4_r2_22.err; tail -1 gpurun_out/bench_c4_r2_22.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_22.json')); print('config 4: value', d['value'], d['unit'], 'ms', d['ms_per_step'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2_launches_r2_22.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bz2_ncu22.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2_launches_r2_22.csv 16 2>&1 | tail -20
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/bench_n1_r2_22.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_r2_22.json')); print('config 2: value', d['value'], d['ms_per_step'])"
