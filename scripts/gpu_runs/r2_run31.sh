# round 2, call 31: BZip2 -- the recorder writes bytes (records only for long runs), RLE1 output in 4 groups with early D2H, slot placement by words
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_gpu.py tests/test_zz_bzip2_damaged_gpu.py tests/test_bzip2_shard.py tests/test_zip_gpu.py tests/test_zz_file_codec_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python scripts/bench_bz2_small.py 2>&1 | tail -4
for eg in 4 1; do
B200Z_BZ2_EMIT_GROUPS=$eg timeout 900 python bench.py --config 4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_r2_31_$eg.json 2> gpurun_out/bench_c4_r2_31_$eg.err; tail -1 gpurun_out/bench_c4_r2_31_$eg.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_31_$eg.json')); print('config 4 emit groups $eg: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bz2_launches_r2_31.csv python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bz2_ncu31.log 2>&1
python scripts/launch_summary.py gpurun_out/bz2_launches_r2_31.csv 40 2>&1 | grep -v "bz2e::" | head -18
