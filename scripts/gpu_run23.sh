mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 90 python scripts/tc_probe.py cin > gpurun_out/probe_cin.log 2>&1; echo "cin probe exit $?"; tail -7 gpurun_out/probe_cin.log | cut -c1-120
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  +assert|passed|failed" gpurun_out/pytest_gpu.log | tail -20 | cut -c1-200
for w in xdeepfm fibinet; do
  timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
  grep '^{' gpurun_out/bench_$w.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][:30], d['ms_per_step'], d['value']); print({k:round(v['ms_per_step'],3) for k,v in sorted(d['per_entry_ms'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]})"
done
timeout -s KILL 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-1800
