# round 2, call 20: the whole GPU suite on the current tree; config 4 with K8 in groups (early D2H)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for gr in 1 4 8; do
B200Z_BZ2_GROUPS=$gr timeout 900 python bench.py --config 4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_r2_20_$gr.json 2> gpurun_out/bench_c4_r2_20_$gr.err; tail -1 gpurun_out/bench_c4_r2_20_$gr.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_r2_20_$gr.json')); print('config 4 groups $gr: value', d['value'], d['unit'], 'ms', d['ms_per_step'])"
done
