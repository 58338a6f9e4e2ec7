mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_round2.py -m gpu -q -x --timeout 600 -p no:cacheprovider --tb=short > gpurun_out/pytest_sub.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_sub.log | cut -c1-300
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_1gpu.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'graph', d['impl_detail']['cuda_graph']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()}); print({k:(v.get('ms_per_step'), v.get('error')) for k,v in d.get('secondary',{}).items()}); f=d['secondary'].get('deepfm_fast_tf32',{}); print({k:f.get(k) for k in ('ms_per_step','value','logit_rel_err_vs_oracle','grad_rel_err_vs_oracle','meets_parity_bar','error')}); print(d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline'])"
