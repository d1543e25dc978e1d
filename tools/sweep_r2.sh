#!/bin/bash
# Round-2 tuning sweeps on one B200 (run under gpurun): class-0 table size x packed load width on the
# alarm-shaped step; class-3 variants and TMA staging on the diabetes-shaped step.  Lines -> gpurun_out/sw_*.json
run() { tag=$1; shift; env "$@" python bench.py --steps 4 --warmup 2 --no-cpu-baseline --stream-dags 0 > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err || echo "FAILED $tag"; }
rund() { tag=$1; shift; env "$@" python bench.py --workload diabetes --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err || echo "FAILED $tag"; }
run base
run v2 BIC_P2_VEC=2
run v1 BIC_P2_VEC=1
run w8k BIC_CLASS0_WORDS=8192
run w8k_v1 BIC_CLASS0_WORDS=8192 BIC_P2_VEC=1
run w12k BIC_CLASS0_WORDS=12288
run w12k_v2 BIC_CLASS0_WORDS=12288 BIC_P2_VEC=2
run w12k_v1 BIC_CLASS0_WORDS=12288 BIC_P2_VEC=1
run w16k BIC_CLASS0_WORDS=16384
run w16k_v1 BIC_CLASS0_WORDS=16384 BIC_P2_VEC=1
rund d_cluster
rund d_passes BIC_CLUSTER=0
rund d_cluster8 BIC_CLUSTER_SIZE=8
rund d_cluster512 BIC_CLUSTER_THREADS=512
rund d_tma BIC_TMA=1
python bench.py --workload pigs --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/sw_pigs.json 2> gpurun_out/sw_pigs.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sw_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'], 3), 'ms', [(c['kernel'][:22], round(c['ms'] / d['steps'], 3), round(c['gbs'])) for c in d['roofline']['classes']])
    except Exception as e:
        print(f, 'ERR', e)
PY
