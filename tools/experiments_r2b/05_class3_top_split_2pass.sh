#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r7_pytest.log
python tools/diag_class3.py > gpurun_out/r7_class3.txt 2>&1; cat gpurun_out/r7_class3.txt
python bench.py --no-cpu-baseline --stream-dags 0 --workload diabetes --steps 10 --warmup 3 > gpurun_out/r7_diabetes.json 2> gpurun_out/r7_diabetes.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r7_diabetes.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'], 4), 'ms', [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
PY
