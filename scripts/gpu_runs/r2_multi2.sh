# round 2: two GPUs -- the library's one-process multi-GPU entry (with the NCCL gather), and bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi2.log
grep -v Warn gpurun_out/pytest_multi2.log | tail -5 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_n2.json'))
    print("N2 value", d['value'], "ms", d['ms_per_step'], "decode_only", d['decode_only'], "strong", d['strong']['value'], d['strong']['ms_per_step'], "e2e", d['e2e']['value'])
    for r in d['per_rank']: print(r)
except Exception as e: print("parse failed", e)
PY
B200Z_FAST=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_n2_fast.json 2> gpurun_out/bench_n2_fast.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n2_fast.json')); print('N2 FAST', d['value'], d['ms_per_step'], d['decode_only'], [ (r['rank'], round(r['ms_per_step'],2), round(r['k_inflate_fast_ms'],2)) for r in d['per_rank']])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 4 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; cut -c1-400 gpurun_out/bench_c4_n2.json; tail -2 gpurun_out/bench_c4_n2.err
B200Z_C5_MEMBERS=512 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config 5 > gpurun_out/bench_c5_n2.json 2> gpurun_out/bench_c5_n2.err; cut -c1-400 gpurun_out/bench_c5_n2.json; tail -2 gpurun_out/bench_c5_n2.err
