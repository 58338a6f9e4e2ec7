mkdir -p gpurun_out; rm -rf gpurun_out/*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
timeout -s KILL 60 ./scripts/probe_umma 2>&1 | tee gpurun_out/probe_umma.log
for w in deepfm xdeepfm fibinet dcn; do
  timeout -s KILL 240 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1
  echo "bench $w exit $?"
  grep '^{' gpurun_out/bench_$w.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][:20], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3)); print({k:round(v['ms_per_step'],3) for k,v in d['per_entry_ms'].items()})"
done
