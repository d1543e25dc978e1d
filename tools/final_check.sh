#!/bin/bash
# Last check of the round at HEAD on one B200: smoke(), the whole GPU suite, the default bench line as the driver runs it
# and the two row-sharded-shaped workloads.
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/chk_smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/chk_smoke.log
python -m pytest tests -m gpu -q > gpurun_out/chk_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/chk_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/chk_alarm.json 2> gpurun_out/chk_alarm.err || echo FAILED alarm
for w in diabetes pigs; do
  python bench.py --no-cpu-baseline --workload $w --steps 10 --warmup 3 > gpurun_out/chk_$w.json 2> gpurun_out/chk_$w.err || echo FAILED $w
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/chk_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d['roofline']['kernel'][:20], d['roofline'].get('frac'), (d.get('stream_1m') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
