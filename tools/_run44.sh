python -m pytest tests -m gpu -x -q > gpurun_out/pytest44.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest44.log
for w in diabetes pigs; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench44_$w.log 2>&1; echo $w rc=$?; done
python bench.py --no-cpu-baseline > gpurun_out/bench44_alarm.log 2>&1; echo alarm rc=$?
ncu --kernel-name 'regex:^(k_|ncclDev)' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_diabetes_launches.csv python bench.py --workload diabetes --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu44.log 2>&1; echo ncu rc=$?
ncu --kernel-name 'regex:^(k_|ncclDev)' --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01j_alarm_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu44a.log 2>&1; echo ncu rc=$?
