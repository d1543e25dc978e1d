#!/bin/bash
# HEAD verification: GPU tests, default line, pigs/diabetes with and without two row groups in flight
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_alarm.json 2> gpurun_out/r1_alarm.err || echo FAILED alarm
B="python bench.py --no-cpu-baseline --steps 10 --warmup 3"
for w in diabetes pigs; do
  $B --workload $w > gpurun_out/r1_${w}_two.json 2> gpurun_out/r1_${w}_two.err || echo FAILED $w
  BIC_U8_TWO=0 BIC_P2_TWO=0 $B --workload $w > gpurun_out/r1_${w}_one.json 2> gpurun_out/r1_${w}_one.err || echo FAILED $w one
  BIC_U8_TWO=3 BIC_P2_TWO=1 $B --workload $w > gpurun_out/r1_${w}_three.json 2> gpurun_out/r1_${w}_three.err || echo FAILED $w three
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r1_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms', round(d['value']), 'DAGs/s  e2e', round(d['e2e']['value']), d.get('roofline',{}).get('achieved'))
    except Exception as e:
        print(f, 'ERR', e)
PY
