mkdir -p gpurun_out
rm -rf gpurun_out/*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.log; nproc >> gpurun_out/gpu.log; lscpu | grep "Model name" >> gpurun_out/gpu.log
timeout -s KILL 120 python scripts/pk_probe.py check > gpurun_out/pk_check.log 2>&1; rc=$?; tail -12 gpurun_out/pk_check.log | cut -c1-200
if [ $rc -eq 0 ]; then echo "PK OK"; else echo "PK FAILED rc=$rc -> tc1"; export CTR_GEMM=tc1; fi
timeout -s KILL 200 python scripts/pk_probe.py bench > gpurun_out/pk_bench.log 2>&1; cat gpurun_out/pk_bench.log | cut -c1-220
timeout -s KILL 200 python scripts/zipf_diag.py > gpurun_out/zipf_diag.log 2>&1; tail -8 gpurun_out/zipf_diag.log | cut -c1-300
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|assert|passed|failed" gpurun_out/pytest_gpu.log | tail -40 | cut -c1-250
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --workload deepfm --no-cpu-baseline > gpurun_out/bench_deepfm.log 2>&1; echo "exit $?" >> gpurun_out/bench_deepfm.log
tail -2 gpurun_out/bench_deepfm.log | cut -c1-3500
for w in dcn xdeepfm fibinet; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "exit $?" >> gpurun_out/bench_$w.log
  tail -2 gpurun_out/bench_$w.log | cut -c1-600
done
for w in deepfm xdeepfm; do
CTR_PROFILE_REGION=1 timeout -s KILL 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/ncu_launch_$w.log 2>&1
echo "ncu launches $w exit $?"
done
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:'gather_fwd_vec|scatter_bwd_vec|gemm_pk|gemm_tc|pack_kvec|pack_trans|plan_insert|plan_finalize' -c 14 -o gpurun_out/full_deepfm -f python scripts/ncu_target.py DeepFM 1 > gpurun_out/ncu_full_deepfm.log 2>&1; echo "ncu full deepfm exit $?"
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:'cin_tc' -c 6 -o gpurun_out/full_xdeepfm -f python scripts/ncu_target.py xDeepFM 1 > gpurun_out/ncu_full_xdeepfm.log 2>&1; echo "ncu full xdeepfm exit $?"
for r in deepfm xdeepfm; do
  ncu -i gpurun_out/full_$r.ncu-rep --page raw --csv > gpurun_out/full_${r}_raw.csv 2>/dev/null
  ncu -i gpurun_out/full_$r.ncu-rep --page source --csv > gpurun_out/full_${r}_source.csv 2>/dev/null
done
du -sm gpurun_out; ls -la gpurun_out
# keep the merge under 64 MiB: drop the largest report(s) if needed (the CSV pages stay)
while [ $(du -sm gpurun_out | cut -f1) -gt 58 ]; do f=$(ls -S gpurun_out/*.ncu-rep gpurun_out/*_source.csv 2>/dev/null | head -1); [ -z "$f" ] && break; echo "dropping $f"; rm -f "$f"; done
du -sm gpurun_out
