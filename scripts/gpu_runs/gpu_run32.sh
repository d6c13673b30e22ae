mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-2500 > gpurun_out/bench_4gpu.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_4gpu.json').read())
print('N=4 value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'allgather', d.get('with_allgather',{}).get('value'))
PY
for n in 1 2 4; do
BZ_MIB=512 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/bench_bz2_multi.py 2>&1 | grep metric | tail -1 | tee gpurun_out/bz2_multi_$n.json
done
