#!/bin/bash
# A/B on one box: final stage of the packed path with the integer dot product (default build) vs PRMT / IMAD / extract
# (libbicgpu_nodp4a.so, -DBIC_P2_DP4A=0).
python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r9_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
NP=$PWD/dags_vae_search_b200/csrc/libbicgpu_nodp4a.so
for i in 1 2; do
  $B --steps 6 --warmup 3 > gpurun_out/r9_alarm_dp4a_$i.json 2>> gpurun_out/r8.err || echo FAILED
  BIC_LIB=$NP $B --steps 6 --warmup 3 > gpurun_out/r9_alarm_nodp4a_$i.json 2>> gpurun_out/r8.err || echo FAILED
  $B --workload pigs --steps 10 --warmup 3 > gpurun_out/r9_pigs_dp4a_$i.json 2>> gpurun_out/r8.err || echo FAILED
  BIC_LIB=$NP $B --workload pigs --steps 10 --warmup 3 > gpurun_out/r9_pigs_nodp4a_$i.json 2>> gpurun_out/r8.err || echo FAILED
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r9_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'), d.get('build', '')[-40:])
    except Exception as e:
        print(f, 'ERR', e)
PY
