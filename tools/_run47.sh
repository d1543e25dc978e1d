python -m pytest tests -m gpu -x -q > gpurun_out/pytest47.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest47.log
for w in diabetes pigs; do python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench47_$w.log 2>&1; echo $w rc=$?; done
for w in asia sachs synthetic_v12_c2; do python bench.py --workload $w --steps 300 --warmup 5 > gpurun_out/bench47_$w.log 2>&1; echo $w rc=$?; done
python bench.py > gpurun_out/bench47_alarm.log 2>&1; echo alarm rc=$?
python bench.py --impl reference > gpurun_out/bench47_alarm_ref.log 2>&1; echo ref rc=$?
