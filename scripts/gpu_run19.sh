mkdir -p gpurun_out
rm -rf gpurun_out/*
timeout -s KILL 120 python scripts/pk_probe.py check > gpurun_out/pk_check.log 2>&1; rc=$?; tail -4 gpurun_out/pk_check.log | cut -c1-200
if [ $rc -eq 0 ]; then echo "PK STREAM OK"; else echo "PK STREAM FAILED rc=$rc -> packed"; export CTR_PK_STREAM=0; fi
timeout -s KILL 200 python scripts/pk_probe.py bench > gpurun_out/pk_bench.log 2>&1; cat gpurun_out/pk_bench.log | cut -c1-220
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|^E  +assert|passed|failed" gpurun_out/pytest_gpu.log | tail -30 | cut -c1-200
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --workload deepfm --no-cpu-baseline > gpurun_out/bench_deepfm.log 2>&1; echo "exit $?" >> gpurun_out/bench_deepfm.log
tail -2 gpurun_out/bench_deepfm.log | cut -c1-300
CTR_PROFILE_REGION=1 timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_deepfm.csv python bench.py --steps 2 --warmup 3 --workload deepfm --no-cpu-baseline > gpurun_out/ncu_launch_deepfm.log 2>&1
echo "ncu launches exit $?"
