python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest43.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest43.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench43_g2.log 2>&1; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload diabetes --steps 20 --warmup 3 > gpurun_out/bench43_g2_diabetes.log 2>&1; echo rc=$?
