python -m pytest tests -m gpu -x -q -k "class3 or slice_choice or large_N or row_sharded" > gpurun_out/pytest41.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest41.log
python bench.py --workload diabetes --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_diabetes.log 2>&1; echo rc=$?
BIC_CLASS2_THREADS=1024 python bench.py --workload diabetes --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench41_diabetes_1024.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01i_diabetes_launches.csv python bench.py --workload diabetes --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu41.log 2>&1; echo ncu rc=$?
