mkdir -p gpurun_out
rm -f gpurun_out/*.log
probe() { timeout -s KILL 90 python scripts/tc_probe.py $1 > gpurun_out/probe_$1.log 2>&1; rc=$?; tail -6 gpurun_out/probe_$1.log | cut -c1-160; return $rc; }
if probe scalar; then echo "TC scalar OK"; else echo "TC FAILED -> simt"; export CTR_GEMM=simt; fi
if probe kvec; then echo "KVEC OK"; else echo "KVEC FAILED -> scalar loads"; export CTR_TC_LOAD=s; fi
if probe mnvec; then echo "TRANS OK"; else echo "TRANS FAILED -> k"; [ -z "$CTR_TC_LOAD" ] && export CTR_TC_LOAD=k; fi
if probe cin; then echo "CIN TC OK"; else echo "CIN TC bwd FAILED -> simt bwd"; export CTR_CIN_TC_BWD=0; fi
env | grep CTR_ > gpurun_out/env_used.log
timeout -s KILL 200 python scripts/tc_timeline.py > gpurun_out/tc_timeline.log 2>&1; cat gpurun_out/tc_timeline.log | cut -c1-200
timeout -s KILL 200 python scripts/zipf_diag.py > gpurun_out/zipf_diag.log 2>&1; tail -6 gpurun_out/zipf_diag.log | cut -c1-250
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -q --timeout 120 -p no:cacheprovider 2>&1 | tail -30 | cut -c1-220 > gpurun_out/pytest_gemm.log; tail -4 gpurun_out/pytest_gemm.log
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
for w in deepfm dcn xdeepfm fibinet; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
  tail -2 gpurun_out/bench_$w.log | cut -c1-200
done
ls -la gpurun_out | head -30
