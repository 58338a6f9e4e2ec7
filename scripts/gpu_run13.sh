mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 120 python scripts/pk_probe.py check > gpurun_out/pk_check.log 2>&1; rc=$?; tail -3 gpurun_out/pk_check.log | cut -c1-200
if [ $rc -eq 0 ]; then echo "PK OK"; else echo "PK FAILED rc=$rc -> tc1"; export CTR_GEMM=tc1; fi
timeout -s KILL 200 python scripts/pk_probe.py bench > gpurun_out/pk_bench.log 2>&1; cat gpurun_out/pk_bench.log | cut -c1-220
timeout -s KILL 200 python scripts/zipf_diag2.py 1.05 > gpurun_out/zipf_diag2.log 2>&1; grep -E "engine|layer" gpurun_out/zipf_diag2.log | cut -c1-400
timeout -s KILL 200 python scripts/zipf_diag2.py 0 > gpurun_out/zipf_diag2_uniform.log 2>&1; grep -E "engine|layer" gpurun_out/zipf_diag2_uniform.log | cut -c1-400
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q --timeout 200 -p no:cacheprovider --tb=short > gpurun_out/pytest_gemm.log 2>&1; tail -3 gpurun_out/pytest_gemm.log | cut -c1-200
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --workload deepfm --no-cpu-baseline > gpurun_out/bench_deepfm.log 2>&1; echo "exit $?" >> gpurun_out/bench_deepfm.log
tail -2 gpurun_out/bench_deepfm.log | cut -c1-400
CTR_PROFILE_REGION=1 timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 2 --warmup 3 --workload deepfm --no-cpu-baseline > gpurun_out/ncu_launch_deepfm.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:'plan_insert|plan_finalize|scatter_bwd|colsum|lin_dense|rowgrad_prep|rowdot_bwd' -c 9 -o gpurun_out/full_misc -f python scripts/ncu_target.py DeepFM 1 > gpurun_out/ncu_full_misc.log 2>&1; echo "ncu full exit $?"
ncu -i gpurun_out/full_misc.ncu-rep --page raw --csv > gpurun_out/full_misc_raw.csv 2>/dev/null
ncu -i gpurun_out/full_misc.ncu-rep --page source --csv > gpurun_out/full_misc_source.csv 2>/dev/null
while [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; do f=$(ls -S gpurun_out/*.ncu-rep gpurun_out/*_source.csv 2>/dev/null | head -1); [ -z "$f" ] && break; echo "dropping $f"; rm -f "$f"; done
du -sm gpurun_out
