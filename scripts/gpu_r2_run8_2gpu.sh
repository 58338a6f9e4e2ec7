mkdir -p gpurun_out; rm -rf gpurun_out/*
nvidia-smi --query-gpu=index,name --format=csv | tail -2
timeout -s KILL 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_round2.py -m gpu -q -k "sharded or cuda1" --timeout 500 -p no:cacheprovider --tb=short > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest exit $?"
grep -E "^(FAILED|ERROR)|passed|failed|SHARDED|worst|rank|Error" gpurun_out/pytest_2gpu.log | tail -30 | cut -c1-400
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu exit $? (124/137 = hang)"
grep -v "^{" gpurun_out/bench_2gpu.log | tail -15 | cut -c1-300
grep '^{' gpurun_out/bench_2gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'graph', d['impl_detail']['cuda_graph'], d['impl_detail']['graph_error']); print(d['parity']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
