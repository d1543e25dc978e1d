#!/bin/bash
# Closing measurements of round 2 (second session) on one B200, run under gpurun: GPU tests, every BASELINE config
# through bench.py at HEAD, the reference arm, the batch-size table, the launch list of the default step and
# ncu --set full of the count kernels (summarised on the box; the .ncu-rep files stay there).  -> gpurun_out/fin_*
python -m pytest tests -m gpu -q > gpurun_out/fin_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/fin_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
python bench.py --steps 20 --warmup 5 > gpurun_out/fin_alarm.json 2> gpurun_out/fin_alarm.err || echo FAILED alarm
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/fin_reference_arm.json 2> gpurun_out/fin_reference_arm.err || echo FAILED reference
for w in asia sachs synthetic_v12_c2 diabetes pigs; do
  python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/fin_$w.json 2> gpurun_out/fin_$w.err || echo FAILED $w
done
for b in 1024 16384 65536; do
  $B --batch $b --steps 4 --warmup 2 > gpurun_out/fin_batch$b.json 2> gpurun_out/fin_batch$b.err || echo FAILED batch $b
done
L="$B --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/fin_alarm_launches.csv $L > gpurun_out/fin_ncu.log 2>&1; echo ncu rc=$?
python tools/ncu_summary.py shares gpurun_out/fin_alarm_launches.csv > gpurun_out/fin_alarm_launch_shares.txt
full() {  # name, kernel regex, count, bench args...
  local name=$1 rx=$2 cnt=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -f -o /tmp/$name $B "$@" > gpurun_out/${name}_ncu.log 2>&1; echo ncu full $name rc=$?
  python tools/ncu_summary.py raw /tmp/$name.ncu-rep > gpurun_out/${name}_ncu_summary.txt
  rm -f /tmp/$name.ncu-rep
}
full fin_alarm_kcount k_count 2 --steps 1 --warmup 1
full fin_diabetes_kcount k_count 4 --workload diabetes --steps 1 --warmup 1
full fin_pigs_kcount k_count 1 --workload pigs --steps 1 --warmup 1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/fin_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value'], 1), d['unit'], 'e2e', round(d['e2e']['value'], 1), [round(c['ms'] / c['launches'], 4) for c in d.get('roofline', {}).get('classes', [])], d.get('cpu_baseline', {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
du -sh gpurun_out
