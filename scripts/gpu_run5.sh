mkdir -p gpurun_out
rm -f gpurun_out/*.log
probe() { timeout -s KILL 90 python scripts/tc_probe.py $1 > gpurun_out/probe_$1.log 2>&1; rc=$?; tail -6 gpurun_out/probe_$1.log | cut -c1-160; return $rc; }
if probe scalar; then echo "TC scalar OK"; else echo "TC FAILED -> simt"; export CTR_GEMM=simt; fi
if probe kvec; then echo "KVEC OK"; else echo "KVEC FAILED -> scalar loads"; export CTR_TC_LOAD=s; fi
if probe mnvec; then echo "TRANS OK"; else echo "TRANS FAILED -> k"; [ -z "$CTR_TC_LOAD" ] && export CTR_TC_LOAD=k; fi
if probe cin; then echo "CIN TC OK"; else echo "CIN TC bwd FAILED -> simt bwd"; export CTR_CIN_TC_BWD=0; fi
env | grep CTR_ > gpurun_out/env_used.log
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -q --timeout 120 -p no:cacheprovider 2>&1 | tail -30 | cut -c1-220 > gpurun_out/pytest_gemm.log; tail -8 gpurun_out/pytest_gemm.log
timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-200
for w in deepfm dcn xdeepfm fibinet; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
  tail -2 gpurun_out/bench_$w.log | cut -c1-200
done
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|gather_fwd_vec|scatter_bwd_vec" --launch-skip 0 -c 8 -o gpurun_out/prof_r1c python scripts/ncu_target.py DeepFM 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
