#!/bin/bash
# Round-2 closing measurements on one B200 (run under gpurun): every BASELINE config through bench.py at HEAD,
# batch-size table, launch list of the default step.  Lines -> gpurun_out/fin_*.json
B="python bench.py --no-cpu-baseline"
python bench.py --steps 20 --warmup 5 > gpurun_out/fin_alarm.json 2> gpurun_out/fin_alarm.err || echo FAILED alarm
for w in asia sachs synthetic_v12_c2 diabetes pigs; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err || echo FAILED $w
done
for b in 1024 16384 65536; do
  $B --batch $b --steps 4 --warmup 2 --stream-dags 0 > gpurun_out/fin_batch$b.json 2> gpurun_out/fin_batch$b.err || echo FAILED batch $b
done
L="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --stream-dags 0"
$L > gpurun_out/fin_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02k_alarm_launches.csv $L > gpurun_out/fin_ncu.log 2>&1; echo ncu rc=$?
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/fin_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms', round(d['value']), 'DAGs/s  e2e', round(d['e2e']['value']), d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
