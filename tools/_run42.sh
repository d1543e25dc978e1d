python -m pytest tests -m gpu -x -q > gpurun_out/pytest42.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest42.log
python bench.py --workload diabetes --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench42_diabetes.log 2>&1; echo rc=$?
ncu --kernel-name 'regex:^(k_|ncclDev)' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_diabetes_launches.csv python bench.py --workload diabetes --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu42.log 2>&1; echo ncu rc=$?
ncu --kernel-name 'regex:^(k_|ncclDev)' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_sachs_launches.csv python bench.py --workload sachs --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu42s.log 2>&1; echo ncu rc=$?
