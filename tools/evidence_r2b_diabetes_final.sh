#!/bin/bash
# Evidence at the final HEAD (after (p)): ncu --set full of the four count kernels of the diabetes-shaped step and
# the launch list of that step; summaries only (the .ncu-rep stays on the box).
B="python bench.py --no-cpu-baseline --stream-dags 0"
L="$B --workload diabetes --steps 2 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r19_diabetes_launches.csv $L > gpurun_out/r19_ncu.log 2>&1; echo ncu rc=$?
python tools/ncu_summary.py shares gpurun_out/r19_diabetes_launches.csv > gpurun_out/r19_diabetes_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:k_count -c 4 -f -o /tmp/r19_diabetes_kcount $B --workload diabetes --steps 1 --warmup 1 > gpurun_out/r19_ncufull.log 2>&1; echo ncu full rc=$?
python tools/ncu_summary.py raw /tmp/r19_diabetes_kcount.ncu-rep > gpurun_out/r19_diabetes_kcount_ncu_summary.txt
head -14 gpurun_out/r19_diabetes_launch_shares.txt
python - <<'PY'
rows=[l.rstrip('\n') for l in open('gpurun_out/r19_diabetes_kcount_ncu_summary.txt')]
for l in rows:
    parts=l.split('|')
    if len(parts)>=5 and any(k in parts[0] for k in ('Kernel','Grid','Block','time_duration','sm__throughput','issue_active','inst_executed.sum','alu_cycles','l1tex__data_pipe_lsu_wavefronts.avg','dram_throughput')):
        print(parts[0][:60].strip().ljust(60), '|', ' | '.join(x.strip()[:28] for x in parts[1:]))
PY
