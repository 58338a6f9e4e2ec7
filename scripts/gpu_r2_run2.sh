mkdir -p gpurun_out; rm -rf gpurun_out/*
nvidia-smi --query-gpu=index,name --format=csv | tail -2
echo "== ts_probe check"; timeout -s KILL 240 python scripts/ts_probe.py check > gpurun_out/ts_check.log 2>&1; echo "ts check exit $?"; tail -25 gpurun_out/ts_check.log | cut -c1-200
echo "== ts_probe bench"; timeout -s KILL 120 python scripts/ts_probe.py bench > gpurun_out/ts_bench.log 2>&1; echo "ts bench exit $?"; tail -6 gpurun_out/ts_bench.log | cut -c1-250
echo "== pytest (TS off)"
CTR_GEMM_TS=0 timeout -s KILL 1300 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -40 | cut -c1-300
echo "== pytest gemm+parity (TS on)"
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q --timeout 600 -p no:cacheprovider --tb=short > gpurun_out/pytest_ts.log 2>&1
echo "pytest TS exit $?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_ts.log | tail -20 | cut -c1-300
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log | cut -c1-200
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?"; tail -c 2500 gpurun_out/bench.log
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -c 1200 gpurun_out/bench_ref.log
