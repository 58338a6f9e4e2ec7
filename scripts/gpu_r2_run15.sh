mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 240 python scripts/ts_probe.py check > gpurun_out/ts_check.log 2>&1; echo "ts check exit $?"
grep -E "TS=1|rel err" gpurun_out/ts_check.log | grep -v "TS=0" | cut -c1-120 | tail -14
timeout -s KILL 120 python scripts/ts_probe.py bench > gpurun_out/ts_bench.log 2>&1; echo "ts bench exit $?"
grep -E "fwd|dgrad" gpurun_out/ts_bench.log | cut -c1-200
timeout -s KILL 200 python scripts/tsw_probe.py check > gpurun_out/tsw_check.log 2>&1; echo "tsw check exit $?"
timeout -s KILL 120 python scripts/tsw_probe.py bench > gpurun_out/tsw_bench.log 2>&1; echo "tsw bench exit $?"
tail -4 gpurun_out/tsw_bench.log | cut -c1-300
timeout -s KILL 120 python scripts/ts_timeline.py fwd1 > gpurun_out/ts_tl_fwd1.log 2>&1
sed -n 1,3p\;14,22p gpurun_out/ts_tl_fwd1.log | cut -c1-130
timeout -s KILL 120 python scripts/tsw_timeline.py dw1 > gpurun_out/tsw_tl_dw1.log 2>&1
sed -n 22,28p\;43,45p gpurun_out/tsw_tl_dw1.log
