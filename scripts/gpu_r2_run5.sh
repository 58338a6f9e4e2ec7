mkdir -p gpurun_out; rm -rf gpurun_out/*
for s in fwd1 fwd2; do timeout -s KILL 120 python scripts/ts_timeline.py $s > gpurun_out/timeline_$s.log 2>&1; echo "timeline $s exit $?"; cat gpurun_out/timeline_$s.log | cut -c1-120; done
