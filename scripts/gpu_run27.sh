mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 200 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 120 -p no:cacheprovider --tb=short -k "tower_large or dnn_layer" > gpurun_out/pytest_tower.log 2>&1; grep -E "^(FAILED|ERROR)|^E  +assert|passed|failed" gpurun_out/pytest_tower.log | tail -8 | cut -c1-200
timeout -s KILL 200 python bench.py --steps 10 --warmup 3 --workload xdeepfm --no-cpu-baseline > gpurun_out/bench_xdeepfm.log 2>&1; echo "exit $?" >> gpurun_out/bench_xdeepfm.log
grep '^{' gpurun_out/bench_xdeepfm.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][:30], d['ms_per_step'], d['value']); print({k:round(v['ms_per_step'],3) for k,v in sorted(d['per_entry_ms'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:6]})"
tail -2 gpurun_out/bench_xdeepfm.log | cut -c1-300
timeout -s KILL 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default bench exit $?"; tail -1 gpurun_out/bench_default.log | cut -c1-2600
