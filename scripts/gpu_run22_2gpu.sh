mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py > gpurun_out/sharded_check.log 2>&1; echo "sharded_check exit $?"; grep -E "rank [01]:|PASSED|FAILED|Error" gpurun_out/sharded_check.log | tail -5 | cut -c1-250
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu exit $?"; grep '^{' gpurun_out/bench_2gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
