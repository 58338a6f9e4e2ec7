mkdir -p gpurun_out; rm -rf gpurun_out/*
CTR_GEMM_TS=1 timeout -s KILL 120 python scripts/ts_timeline.py fwd1 > gpurun_out/ts_tl_fwd1.log 2>&1
sed -n 1,3p\;12,22p\;30,33p gpurun_out/ts_tl_fwd1.log | cut -c1-130
CTR_GEMM_TS=1 timeout -s KILL 120 python scripts/ts_timeline.py dx1 > gpurun_out/ts_tl_dx1.log 2>&1
sed -n 1,3p\;8,14p\;18,22p gpurun_out/ts_tl_dx1.log | cut -c1-130
CTR_GEMM_TS=1 timeout -s KILL 120 python scripts/ts_probe.py bench 2>&1 | grep -E "fwd|dgrad" | cut -c1-200
