#!/bin/bash
# After k_derive (warp per target cell for wide sums) and k_acyclic (register / shared-memory resident masks):
# GPU tests, then the steps those kernels matter for.
python -m pytest tests -m gpu -x -q > gpurun_out/r5_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r5_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
for w in diabetes pigs sachs asia synthetic_v12_c2; do
  $B --workload $w --steps 10 --warmup 3 > gpurun_out/r5_$w.json 2> gpurun_out/r5_$w.err || echo FAILED $w
done
for w in diabetes sachs asia; do
  L="$B --workload $w --steps 2 --warmup 1"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r5_${w}_launches.csv $L > gpurun_out/r5_ncu_$w.log 2>&1; echo ncu $w rc=$?
  python tools/ncu_summary.py shares gpurun_out/r5_${w}_launches.csv > gpurun_out/r5_${w}_launch_shares.txt
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r5_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), round(d['e2e']['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
head -12 gpurun_out/r5_*_launch_shares.txt
