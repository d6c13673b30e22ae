# round 2, call 21: k_inflate_fast with LZ77 by blocks (default) against the per-byte pass (variant lzold); parity; phase clocks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inflate_gpu.py tests/test_zz_gzip_stream_semantics_gpu.py tests/test_zip_gpu.py -x -q -m gpu 2>&1 | tail -2
for v in "" lzold ""; do
  if [ -n "$v" ]; then export B200Z_LIB=archive_b200/variants/libb200z_$v.so; else unset B200Z_LIB; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side-configs > gpurun_out/var_$v.json 2> gpurun_out/var_$v.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/var_$v.json')); print('variant [$v]: value', round(d['value'],2), round(d['ms_per_step'],3))" || tail -3 gpurun_out/var_$v.err
done
B200Z_LIB=archive_b200/variants/libb200z_prof.so timeout 300 python scripts/fast_prof.py 2>&1 | tail -22
