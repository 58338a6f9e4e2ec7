mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/shard_gather_diag.py > gpurun_out/shard_gather_diag.log 2>&1; echo "diag exit $?"; grep -E "^rank 0|Error|error" gpurun_out/shard_gather_diag.log | head -30 | cut -c1-200
