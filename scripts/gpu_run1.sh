mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_rowwise_plan_properties_at_full_batch 2>&1 | tail -60 > gpurun_out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_deepfm.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_deepfm.log
timeout 300 python bench.py --steps 10 --warmup 3 --workload dcn --no-cpu-baseline > gpurun_out/bench_dcn.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_deepfm.log
