mkdir -p gpurun_out; rm -rf gpurun_out/*
for i in 1 2; do
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/bench_$i.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/bench_$i.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value']); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
