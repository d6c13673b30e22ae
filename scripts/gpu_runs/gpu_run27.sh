mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_zip_gpu.py -x -q > gpurun_out/pytest_zip.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_zip.log
grep -v Warn gpurun_out/pytest_zip.log | tail -3 | cut -c1-250
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_inflate_ -s 6 -c 2 -o gpurun_out/r1_inflate_final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_inflate_final.log 2>&1
ls -la gpurun_out/r1_inflate_final.ncu-rep
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device='cuda')
for name, fn in (('H2D', lambda: d.copy_(h, non_blocking=True)), ('D2H', lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); print(name, 'GB/s', round(3 * n / (time.perf_counter() - t) / 1e9, 1))
PY
