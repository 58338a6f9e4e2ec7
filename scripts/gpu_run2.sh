mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_gpu_gemm.py -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "gemm pytest exit $?" >> gpurun_out/pytest_gemm.log
tail -5 gpurun_out/pytest_gemm.log
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for w in deepfm dcn xdeepfm fibinet; do
  timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
done
CTR_GEMM=simt timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --workload deepfm --no-cpu-baseline > gpurun_out/bench_deepfm_simt.log 2>&1
# launch list of one DeepFM bench (cold-cache, serialised: compare shares)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# full capture of the HBM-bound kernels and the tensor-core GEMM
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gather_fwd_vec|scatter_bwd_vec|gemm_tc_kernel|plan_insert" -s 12 -c 8 -o gpurun_out/prof_r1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
