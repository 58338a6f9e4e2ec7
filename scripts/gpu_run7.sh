mkdir -p gpurun_out
timeout -s KILL 200 python scripts/tc_timeline.py > gpurun_out/tc_timeline.log 2>&1
cat gpurun_out/tc_timeline.log
