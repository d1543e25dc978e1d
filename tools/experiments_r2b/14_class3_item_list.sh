#!/bin/bash
# Class-3 grid with one item per (family, pass it needs) instead of njobs x (most passes): tests, diabetes-shaped step
# with the list / without (BIC_NO_LIST3=1) / with 16-bit counters on top; sachs and asia once more (narrow class 0 again).
python -m pytest tests -m gpu -x -q > gpurun_out/r18_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/r18_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --workload diabetes --steps 10 --warmup 3 > gpurun_out/r18_diabetes_$tag.json 2>> gpurun_out/r18.err || echo "FAILED $tag"; }
run list
run nolist BIC_NO_LIST3=1
run list_u16 BIC_C3_U16=1
run list2
run nolist2 BIC_NO_LIST3=1
run list_u16_2 BIC_C3_U16=1
for w in sachs asia; do
  $B --workload $w --steps 10 --warmup 3 > gpurun_out/r18_$w.json 2>> gpurun_out/r18.err || echo FAILED $w
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r18_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], [round(c['gbs']) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
