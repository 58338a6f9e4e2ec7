mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 200 python scripts/tsw_probe.py check > gpurun_out/tsw_check.log 2>&1; echo "tsw check exit $?"
tail -20 gpurun_out/tsw_check.log | cut -c1-300
timeout -s KILL 120 python scripts/tsw_probe.py bench > gpurun_out/tsw_bench.log 2>&1; echo "tsw bench exit $?"
tail -8 gpurun_out/tsw_bench.log | cut -c1-300
