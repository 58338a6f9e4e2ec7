mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 120 python scripts/ss_timeline.py fwd1 > gpurun_out/ss_tl_fwd1.log 2>&1
cat gpurun_out/ss_tl_fwd1.log | cut -c1-130
timeout -s KILL 120 python scripts/ss_timeline.py dx1 > gpurun_out/ss_tl_dx1.log 2>&1
sed -n 1,3p\;8,20p gpurun_out/ss_tl_dx1.log | cut -c1-130
