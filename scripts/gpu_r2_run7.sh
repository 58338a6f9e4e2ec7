mkdir -p gpurun_out; rm -rf gpurun_out/*
for cfg in "1 0" "16 0" "64 0" "16 1"; do
  set -- $cfg
  echo "== ts bench BREP=$1 BRAW=$2"
  CTR_TS_BREP=$1 CTR_TS_BRAW=$2 timeout -s KILL 120 python scripts/ts_probe.py bench 2>&1 | grep -E "fwd|dgrad" | sed 's/(.*ceiling)  TS=0.*//' | cut -c1-110
done
