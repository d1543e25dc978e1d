#!/bin/bash
# After the nibble pairs + IDP.4A (ALU pipe relieved): which resource binds now?  Table sizes, sorted rows, ncu.
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --steps 4 --warmup 2 > gpurun_out/r10_$tag.json 2>> gpurun_out/r10.err || echo "FAILED $tag"; }
run base
run w8k BIC_CLASS0_WORDS=8192
run w16k BIC_CLASS0_WORDS=16384
run sortlex BENCH_SORT_ROWS=lex
run sortlex_w16k BENCH_SORT_ROWS=lex BIC_CLASS0_WORDS=16384
run c512 BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r10_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
name=r10_alarm_kcount
ncu --set full --clock-control none --import-source on -k regex:k_count -c 2 -f -o /tmp/$name $B --steps 1 --warmup 1 > gpurun_out/${name}_ncu.log 2>&1; echo ncu full rc=$?
python tools/ncu_summary.py raw /tmp/$name.ncu-rep > gpurun_out/${name}_ncu_summary.txt
ncu -i /tmp/$name.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/${name}_details.csv.gz
cut -c1-160 gpurun_out/${name}_ncu_summary.txt
