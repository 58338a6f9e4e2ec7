mkdir -p gpurun_out; rm -rf gpurun_out/*
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_(tsw|pk)_kernel' -s 2 -c 1 -o gpurun_out/tsw_dw1 python scripts/ncu_tsw.py dw1 > gpurun_out/ncu_tsw.log 2>&1; echo "ncu tsw exit $?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_(tsw|pk)_kernel' -s 5 -c 1 -o gpurun_out/ss_dw1 python scripts/ncu_tsw.py dw1 > gpurun_out/ncu_ss.log 2>&1; echo "ncu ss exit $?"
tail -3 gpurun_out/ncu_tsw.log; ls -la gpurun_out
