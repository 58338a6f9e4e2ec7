mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 240 python scripts/pp_probe.py check > gpurun_out/pp_check.log 2>&1; echo "check exit $?"
grep -E "rel err|rror|Traceback|timed out" gpurun_out/pp_check.log | cut -c1-120 | tail -32
timeout -s KILL 120 python scripts/pp_probe.py bench > gpurun_out/pp_bench.log 2>&1; echo "bench exit $?"
grep -E "fwd|dgrad" gpurun_out/pp_bench.log | cut -c1-220
