mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 120 python scripts/tsw_timeline.py dw1 > gpurun_out/tsw_tl_dw1.log 2>&1; echo "exit $?"
timeout -s KILL 120 python scripts/tsw_timeline.py dw2 > gpurun_out/tsw_tl_dw2.log 2>&1; echo "exit $?"
cat gpurun_out/tsw_tl_dw1.log | head -50
