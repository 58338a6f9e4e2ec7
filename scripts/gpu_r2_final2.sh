mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_default.log > gpurun_out/bench_default.json
python -c "
import sys,json
d=json.loads(open('gpurun_out/bench_default.json').read()); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'graph', d['impl_detail']['cuda_graph'], 'launches', d['gpu_launches']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()}); print({k:(v.get('ms_per_step'), v.get('error')) for k,v in d.get('secondary',{}).items()}); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('frac_of_3xtf32_ceiling')); print(d['cpu_baseline']['value'])"
