python -m pytest tests -m gpu -x -q > gpurun_out/pytest48.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest48.log
python tools/diag_latency.py > gpurun_out/latency48_fast.log 2>&1; cat gpurun_out/latency48_fast.log
BIC_NO_FAST_SMALL=1 python tools/diag_latency.py > gpurun_out/latency48_general.log 2>&1; cat gpurun_out/latency48_general.log
