#!/bin/bash
# Experiment: dataset rows in lexicographic order (BENCH_SORT_ROWS) -> more lanes of a warp atomic meet in the same
# cell.  Same counts, same scores (checksum must not move).
B="python bench.py --no-cpu-baseline --stream-dags 0"
for m in "" lex card; do
  tag=${m:-none}
  BENCH_SORT_ROWS=$m $B --steps 6 --warmup 3 > gpurun_out/r3_alarm_$tag.json 2> gpurun_out/r3_alarm_$tag.err || echo FAILED alarm $tag
  BENCH_SORT_ROWS=$m $B --workload pigs --steps 10 --warmup 3 > gpurun_out/r3_pigs_$tag.json 2> gpurun_out/r3_pigs_$tag.err || echo FAILED pigs $tag
  BENCH_SORT_ROWS=$m $B --workload diabetes --steps 10 --warmup 3 > gpurun_out/r3_diabetes_$tag.json 2> gpurun_out/r3_diabetes_$tag.err || echo FAILED diabetes $tag
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r3_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
