#!/bin/bash
# Bank swizzle of un-replicated packed-path tables: tests, alarm-shaped step with and without.
python -m pytest tests -m gpu -x -q > gpurun_out/r13_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r13_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --steps 4 --warmup 2 > gpurun_out/r13_$tag.json 2>> gpurun_out/r13.err || echo "FAILED $tag"; }
run swz
run noswz BIC_SWIZZLE=0
run swz2
run noswz2 BIC_SWIZZLE=0
$B --workload pigs --steps 10 --warmup 3 > gpurun_out/r13_pigs.json 2>> gpurun_out/r13.err || echo FAILED pigs
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r13_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), round(d['e2e']['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
