mkdir -p gpurun_out; rm -rf gpurun_out/*
echo "== check BRAW(dedicated warps)"; timeout -s KILL 240 python scripts/ts_probe.py check 2>&1 | grep -E "TS=1 rel|y rel" | cut -c1-100
for cfg in "1" "0"; do
  echo "== ts bench BRAW=$cfg"
  CTR_TS_BRAW=$cfg timeout -s KILL 120 python scripts/ts_probe.py bench 2>&1 | grep -E "fwd|dgrad" | cut -c1-200
done
timeout -s KILL 120 python scripts/ts_timeline.py fwd1 2>&1 | sed -n 1,3p\;12,20p | cut -c1-120
