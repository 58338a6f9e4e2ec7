mkdir -p gpurun_out
rm -rf gpurun_out/*
nvidia-smi -L > gpurun_out/gpus.log
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py > gpurun_out/sharded_check.log 2>&1; echo "sharded_check exit $?"; grep -v "^\s*$" gpurun_out/sharded_check.log | tail -12 | cut -c1-250
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench 2gpu exit $?"; grep -v "^\s*$" gpurun_out/bench_2gpu.log | tail -3 | cut -c1-1800
for spw in 2 1; do
CTR_SCATTER_SPW=$spw timeout -s KILL 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_spw$spw.log 2>&1; echo "spw=$spw"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_1gpu_spw$spw.log | head -1; grep -o '"ctr_scatter_bwd_rowwise": {[^}]*}' gpurun_out/bench_1gpu_spw$spw.log | head -2
done
