mkdir -p gpurun_out
rm -f gpurun_out/*.log
if timeout -s KILL 120 python scripts/tc_probe.py > gpurun_out/tc_probe.log 2>&1; then echo "TC OK"; else echo "TC FAILED -> simt"; export CTR_GEMM=simt; fi
cat gpurun_out/tc_probe.log | tail -6
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -q --timeout 120 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_gemm.log
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q --timeout 100 -p no:cacheprovider -k "xdeepfm_small or cin" 2>&1 | tail -15 > gpurun_out/pytest_cin_first.log
if grep -q "passed" gpurun_out/pytest_cin_first.log && ! grep -q "failed" gpurun_out/pytest_cin_first.log; then echo "CIN TC OK"; else echo "CIN TC suspicious"; tail -5 gpurun_out/pytest_cin_first.log; fi
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-200
for w in deepfm dcn xdeepfm fibinet; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
  tail -2 gpurun_out/bench_$w.log | cut -c1-300
done
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload deepfm --no-cpu-baseline --no-graph > gpurun_out/bench_deepfm_nograph.log 2>&1
CTR_GEMM=simt timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload deepfm --no-cpu-baseline > gpurun_out/bench_deepfm_simt.log 2>&1
# ncu: launch list (my kernels only) + full capture of the top kernels, on the light target
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:_kernel -c 200 --csv --log-file gpurun_out/launches_deepfm.csv python scripts/ncu_target.py DeepFM 2 > gpurun_out/ncu_launches.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"gather_fwd_vec|scatter_bwd_vec|gemm_tc_kernel|plan_insert" --launch-skip 8 -c 6 -o gpurun_out/prof_r1 python scripts/ncu_target.py DeepFM 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
