#!/bin/bash
# Class 3 with sub-ranges along the first parent's states (top split): parity tests, then the diabetes-shaped step
# with and without it.
python -m pytest tests -m gpu -x -q > gpurun_out/r6_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r6_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
for v in 1 0 1 0; do
  BIC_TOPSPLIT=$v $B --workload diabetes --steps 10 --warmup 3 > gpurun_out/r6_diabetes_top${v}_$RANDOM.json 2>> gpurun_out/r6_diabetes.err || echo FAILED $v
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r6_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), round(d['e2e']['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], [round(c['gbs']) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
