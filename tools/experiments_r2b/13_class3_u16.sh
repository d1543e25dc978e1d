#!/bin/bash
# Class 3 with 16-bit counters (BIC_C3_U16=1) and the wide class-0 shape on the uint8 path: tests, diabetes-shaped step.
python -m pytest tests -m gpu -x -q > gpurun_out/r17_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r17_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
for i in 1 2; do
  for v in 0 1; do
    BIC_C3_U16=$v $B --workload diabetes --steps 10 --warmup 3 > gpurun_out/r17_diabetes_u16_${v}_$i.json 2>> gpurun_out/r17.err || echo FAILED $v
  done
done
for w in asia sachs synthetic_v12_c2; do
  $B --workload $w --steps 10 --warmup 3 > gpurun_out/r17_$w.json 2>> gpurun_out/r17.err || echo FAILED $w
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r17_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], [round(c['gbs']) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
