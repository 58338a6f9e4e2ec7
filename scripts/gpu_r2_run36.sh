mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 300 python -m pytest tests/test_activation.py tests/test_gpu_fit.py -q --timeout 300 -p no:cacheprovider --tb=short > gpurun_out/pytest_act.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/pytest_act.log | cut -c1-300
