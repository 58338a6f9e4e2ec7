mkdir -p gpurun_out; rm -rf gpurun_out/*
nvidia-smi --query-gpu=index,name --format=csv | tail -8
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --allow-eager --no-secondary > gpurun_out/bench_8gpu.log 2>&1; echo "bench 8gpu exit $? (124/137 = hang)"
grep -v "^{" gpurun_out/bench_8gpu.log | grep -v "^\*\|OMP_NUM\|warn" | tail -25 | cut -c1-400
grep '^{' gpurun_out/bench_8gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'graph', d['impl_detail']['cuda_graph'], d['impl_detail']['graph_error']); print(d['parity']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
