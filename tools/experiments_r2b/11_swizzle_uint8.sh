#!/bin/bash
# Bank swizzle on the uint8 path too: tests, diabetes-shaped / small workloads with and without.
python -m pytest tests -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r14_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
for w in diabetes synthetic_v12_c2 sachs asia; do
  for v in 1 0; do
    BIC_SWIZZLE=$v $B --workload $w --steps 10 --warmup 3 > gpurun_out/r14_${w}_swz$v.json 2>> gpurun_out/r14.err || echo FAILED $w $v
  done
done
$B --steps 4 --warmup 2 > gpurun_out/r14_alarm.json 2>> gpurun_out/r14.err || echo FAILED alarm
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r14_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), round(d['e2e']['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
