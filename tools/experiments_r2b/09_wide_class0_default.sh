#!/bin/bash
# New class-0 default (512 threads x 96 KB for mid-size packed tables): tests, alarm / pigs / v12 / batch sizes, ncu.
python -m pytest tests -m gpu -x -q > gpurun_out/r12_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r12_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --steps 4 --warmup 2 > gpurun_out/r12_$tag.json 2>> gpurun_out/r12.err || echo "FAILED $tag"; }
run alarm
run alarm_narrow BIC_CLASS0_WIDE=0
for w in pigs synthetic_v12_c2 asia sachs; do
  $B --workload $w --steps 10 --warmup 3 > gpurun_out/r12_$w.json 2>> gpurun_out/r12.err || echo FAILED $w
done
BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576 $B --workload pigs --steps 10 --warmup 3 > gpurun_out/r12_pigs_wide.json 2>> gpurun_out/r12.err || echo FAILED pigs wide
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r12_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), round(d['e2e']['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
name=r12_alarm_kcount
ncu --set full --clock-control none --import-source on -k regex:k_count -c 2 -f -o /tmp/$name $B --steps 1 --warmup 1 > gpurun_out/${name}_ncu.log 2>&1; echo ncu full rc=$?
python tools/ncu_summary.py raw /tmp/$name.ncu-rep > gpurun_out/${name}_ncu_summary.txt
ncu -i /tmp/$name.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/${name}_details.csv.gz
cut -c1-160 gpurun_out/${name}_ncu_summary.txt
