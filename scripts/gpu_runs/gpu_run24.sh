mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bzip2_shard.py tests/test_bzip2_gpu.py -x -q -m gpu > gpurun_out/pytest_bz2s.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_bz2s.log
grep -v Warn gpurun_out/pytest_bz2s.log | tail -4 | cut -c1-250
BZ_MIB=512 timeout 600 python scripts/bench_bz2_multi.py 2>&1 | tail -1 | tee gpurun_out/bz2_multi_1.json
BZ_MIB=512 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_bz2_multi.py 2>&1 | tail -2 | tee gpurun_out/bz2_multi_2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench_2gpu.json
