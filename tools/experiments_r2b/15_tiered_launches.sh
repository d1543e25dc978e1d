#!/bin/bash
# Tiers: the class-0 / class-1 lists counted by two launches with different CTA shapes (BIC_TIER0 / BIC_TIER1 = cell
# limit of the first launch; 0 = one launch).  Parity with the tiers on, then the alarm-shaped step.
BIC_TIER0=768 BIC_TIER1=3072 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or alarm or derived" > gpurun_out/r20_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r20_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --steps 4 --warmup 2 > gpurun_out/r20_$tag.json 2>> gpurun_out/r20.err || echo "FAILED $tag"; }
run t0_0_t1_0
run t0_768_t1_0 BIC_TIER0=768
run t0_768_t1_3072 BIC_TIER0=768 BIC_TIER1=3072
run t0_0_t1_3072 BIC_TIER1=3072
run t0_1536_t1_3072 BIC_TIER0=1536 BIC_TIER1=3072
run t0_1024_t1_3072 BIC_TIER0=1024 BIC_TIER1=3072
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r20_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
