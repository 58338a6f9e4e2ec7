mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 120 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 100 -p no:cacheprovider --tb=short -k "tower_large" > gpurun_out/pytest_tower.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +assert|passed|failed" gpurun_out/pytest_tower.log | tail -6 | cut -c1-200
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu exit $? (124/137 = hang)"; grep '^{' gpurun_out/bench_2gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config'].get('cuda_graph'))"
