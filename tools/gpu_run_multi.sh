#!/bin/bash
# N-GPU check at HEAD (N = number of visible GPUs): multi-GPU tests, then the bench line with parity_check,
# scaling_context, row_sharded and stream_1m.
N=$(python -c "import torch; print(torch.cuda.device_count())")
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/rm_pytest_${N}gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/rm_pytest_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/rm_bench_${N}gpu.json 2> gpurun_out/rm_bench_${N}gpu.err; echo bench rc=$?
tail -c 1500 gpurun_out/rm_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/rm_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), 'DAGs/s', round(d['ms_per_step'], 3), 'ms e2e', round(d['e2e']['value']))
print('parity_check', d.get('parity_check'))
print('row_sharded', json.dumps(d.get('row_sharded'))[:1500])
print('scaling_context', json.dumps(d.get('scaling_context'))[:800])
print('stream_1m', json.dumps(d.get('stream_1m'))[:600])
PY
