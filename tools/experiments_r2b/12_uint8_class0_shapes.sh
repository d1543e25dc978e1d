#!/bin/bash
# uint8 path (diabetes-shaped): class-0 CTA shape x table size again, now with two row groups in flight.
B="python bench.py --no-cpu-baseline --stream-dags 0 --workload diabetes --steps 10 --warmup 3"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r16_$tag.json 2>> gpurun_out/r16.err || echo "FAILED $tag"; }
run base
run w8k BIC_CLASS0_WORDS=8192
run w12k BIC_CLASS0_WORDS=12288
run c512_w12k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=12288
run c512_w16k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=16384
run c512_w24k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576
run c1024_w24k BIC_CLASS0_THREADS=1024 BIC_CLASS0_WORDS=24576
run c1024_w48k BIC_CLASS0_THREADS=1024 BIC_CLASS0_WORDS=49152
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r16_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
