mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 200 python scripts/tsw_probe.py check > gpurun_out/tsw_check.log 2>&1; echo "tsw check exit $?"
grep -E "worst|rror" gpurun_out/tsw_check.log | cut -c1-200 | tail -3
timeout -s KILL 120 python scripts/tsw_probe.py bench > gpurun_out/tsw_bench.log 2>&1; echo "tsw bench exit $?"
tail -4 gpurun_out/tsw_bench.log | cut -c1-300
timeout -s KILL 120 python scripts/tsw_timeline.py dw1 > gpurun_out/tsw_tl_dw1.log 2>&1
sed -n 22,26p\;43,45p gpurun_out/tsw_tl_dw1.log
timeout -s KILL 120 python scripts/tsw_timeline.py dw2 > gpurun_out/tsw_tl_dw2.log 2>&1
sed -n 12,16p\;43,45p gpurun_out/tsw_tl_dw2.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_1gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'graph', d['impl_detail']['cuda_graph']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()}); print({k:(v.get('ms_per_step'), v.get('error')) for k,v in d.get('secondary',{}).items()}); print(d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_1gpu.log | cut -c1-300
