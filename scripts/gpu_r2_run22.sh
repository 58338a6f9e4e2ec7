mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 240 python scripts/ts_probe.py check > gpurun_out/ts_check.log 2>&1; echo "check exit $?"
grep -E "TS=0|tanh" gpurun_out/ts_check.log | cut -c1-120 | tail -14
timeout -s KILL 120 python scripts/ts_probe.py bench > gpurun_out/ts_bench.log 2>&1
grep -E "fwd|dgrad" gpurun_out/ts_bench.log | cut -c1-220
CTR_PK_ATMA=0 timeout -s KILL 120 python scripts/ts_probe.py bench 2>&1 | grep -E "fwd|dgrad" | cut -c100-220
timeout -s KILL 120 python scripts/ss_timeline.py fwd1 > gpurun_out/ss_tl_fwd1.log 2>&1
sed -n 1,3p\;10,20p\;30,31p gpurun_out/ss_tl_fwd1.log | cut -c1-130
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x --timeout 300 -p no:cacheprovider --tb=short > gpurun_out/pytest_gemm.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gemm.log | cut -c1-300
