#!/bin/bash
# Replica reach after the 16-bit offset cap is lifted on the packed path: CTA shape x table size, sorted rows.
python -m pytest tests -m gpu -x -q > gpurun_out/r11_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r11_pytest.log
B="python bench.py --no-cpu-baseline --stream-dags 0"
run() { tag=$1; shift; env "$@" $B --steps 4 --warmup 2 > gpurun_out/r11_$tag.json 2>> gpurun_out/r11.err || echo "FAILED $tag"; }
run base
run w16k BIC_CLASS0_WORDS=16384
run c512_w24k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576
run c512_w28k BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=28672
run c1024_w48k BIC_CLASS0_THREADS=1024 BIC_CLASS0_WORDS=49152
run sortlex_c512_w24k BENCH_SORT_ROWS=lex BIC_CLASS0_THREADS=512 BIC_CLASS0_WORDS=24576
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r11_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 4), 'ms', round(d['value']), [round(c['ms'] / c['launches'], 4) for c in d['roofline']['classes']], d.get('checksum'))
    except Exception as e:
        print(f, 'ERR', e)
PY
