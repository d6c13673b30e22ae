# round 2, call 10: kernel build variants (helper warps, look-up tries, copy step) on config 2, and the e2e walk both ways
mkdir -p gpurun_out
for v in "" xt64 xt128 t1 t3 s16 xt128t3; do
  if [ -n "$v" ]; then export B200Z_LIB=archive_b200/variants/libb200z_$v.so; else unset B200Z_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/var_$v.json')); print('variant [$v]: value', d['value'], d['ms_per_step'])" || tail -3 gpurun_out/var_$v.err
done
unset B200Z_LIB
for w in 1 0; do
B200Z_GZIP_PIPED_WALK=$w timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-side-configs > gpurun_out/bench_10_w$w.json 2> gpurun_out/bench_10_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/bench_10_w$w.json')); print('piped walk=$w: e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'value', d['value'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
