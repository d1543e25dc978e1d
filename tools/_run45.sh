ncu --kernel-name 'regex:^k_count' --launch-skip 8 --launch-count 4 --set full --clock-control none --import-source on -o gpurun_out/r01j_diabetes_kcount python bench.py --workload diabetes --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu45.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/*.ncu-rep | tail -2
