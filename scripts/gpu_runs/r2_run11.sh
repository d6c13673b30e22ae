# round 2, call 11: where a unit's wall time goes in k_inflate_fast (FP_PROF builds), then the t1 variant again
mkdir -p gpurun_out
for v in prof profxt128; do
  B200Z_LIB=archive_b200/variants/libb200z_$v.so timeout 300 python scripts/fast_prof.py 2>&1 | tail -14
done
