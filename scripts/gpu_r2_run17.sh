mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 200 python scripts/tsw_probe.py check > gpurun_out/tsw_check.log 2>&1; echo "tsw check exit $?"
grep -E "TSW=1|worst" gpurun_out/tsw_check.log | cut -c1-200
timeout -s KILL 120 python scripts/tsw_probe.py bench > gpurun_out/tsw_bench.log 2>&1; echo "tsw bench exit $?"
tail -4 gpurun_out/tsw_bench.log | cut -c1-300
timeout -s KILL 120 python scripts/tsw_timeline.py dw1 > gpurun_out/tsw_tl_dw1.log 2>&1
sed -n 1,3p\;22,28p\;43,45p gpurun_out/tsw_tl_dw1.log
timeout -s KILL 120 python scripts/tsw_timeline.py dw2 > gpurun_out/tsw_tl_dw2.log 2>&1
sed -n 12,16p\;43,45p gpurun_out/tsw_tl_dw2.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
