mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider --tb=short -k "xdeepfm or xDeepFM or cin" > gpurun_out/pytest_sub.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_sub.log | cut -c1-300
for wl in xdeepfm; do
timeout -s KILL 400 python bench.py --workload $wl --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_$wl.log 2>&1; echo "bench $wl exit $?"
grep '^{' gpurun_out/bench_$wl.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'value', d['value'], d['parity']['max_rel_err']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
done
CTR_PROFILE_REGION=1 timeout -s KILL 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_xdeepfm.csv python bench.py --workload xdeepfm --steps 2 --warmup 3 --no-secondary > gpurun_out/bench_ncu_xdeepfm.log 2>&1; echo "ncu exit $?"
grep -E "cin_" gpurun_out/launches_xdeepfm.csv | awk -F'","' '{print $5, $(NF)}' | sort | uniq -c | head -30
