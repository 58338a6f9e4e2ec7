mkdir -p gpurun_out; rm -rf gpurun_out/*
echo "== ts check (cluster 2)"; timeout -s KILL 240 python scripts/ts_probe.py check > gpurun_out/ts_check.log 2>&1; echo "ts check exit $?"; grep -c "rel err" gpurun_out/ts_check.log; grep "TS=1" gpurun_out/ts_check.log | head -8 | cut -c1-120; tail -3 gpurun_out/ts_check.log | cut -c1-200
for cfg in "1 0" "2 0" "4 0" "2 1" "4 1"; do
  set -- $cfg
  echo "== ts bench CLUSTER=$1 RAWD=$2"
  CTR_TS_CLUSTER=$1 CTR_TS_RAWD=$2 timeout -s KILL 120 python scripts/ts_probe.py bench 2>&1 | grep -E "fwd|dgrad" | sed 's/(.*ceiling)  TS=0.*//' | cut -c1-110
done
CTR_TS_CLUSTER=4 timeout -s KILL 120 python scripts/ts_timeline.py fwd1 2>&1 | sed -n 1,3p\;12,22p | cut -c1-120
echo "== cin_v2 tests"; timeout -s KILL 300 python -m pytest tests/test_gpu_cin_v2.py -m gpu -q --timeout 200 -p no:cacheprovider --tb=line 2>&1 | tail -6 | cut -c1-250
