#!/bin/bash
# Round-2 sweep 2: class-0 CTA shape x table size on the alarm-shaped step (all-packed dataset).
run() { tag=$1; shift; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline --stream-dags 0 > gpurun_out/sx_$tag.json 2> gpurun_out/sx_$tag.err || echo "FAILED $tag"; }
run t256_w12k
run t256_w10k BIC_CLASS0_WORDS=10240
run t256_w14k BIC_CLASS0_WORDS=14336
run t512_w12k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=12288
run t512_w16k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=16384
run t512_w24k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576
run t512_w28k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=28672
run t1024_w32k BIC_CLASS0_THREADS=1024 BIC_CLASS0_WORDS=32768
run t1024_w49k BIC_CLASS0_THREADS=1024 BIC_CLASS0_WORDS=49152
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sx_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms', [(c['kernel'][:22], round(c['ms'] / d['steps'], 3), round(c['gbs'])) for c in d['roofline']['classes']])
    except Exception as e:
        print(f, 'ERR', e)
PY
