mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu exit $? (124/137 = hang)"
grep -v "^{" gpurun_out/bench_2gpu.log | grep -v "^\*\|OMP_NUM\|warn\|Warn" | tail -8 | cut -c1-300
grep '^{' gpurun_out/bench_2gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['mode'][:40], 'graph', d['impl_detail']['cuda_graph']); print(d['parity']['ok'], {k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
