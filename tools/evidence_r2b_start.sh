#!/bin/bash
# Round-2 (second session) evidence run on one B200: launch lists of the diabetes / asia / sachs / pigs steps,
# ncu --set full of the count kernels of the alarm- and diabetes-shaped steps at HEAD.  The .ncu-rep files are
# summarised on the box (tools/ncu_summary.py raw) and removed: gpurun_out/ travels back only below 64 MiB.
B="python bench.py --no-cpu-baseline --stream-dags 0"
for w in diabetes asia sachs pigs; do
  L="$B --workload $w --steps 2 --warmup 1"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/r2_${w}_launches.csv $L > gpurun_out/r2_ncu_$w.log 2>&1; echo ncu $w rc=$?
  python tools/ncu_summary.py shares gpurun_out/r2_${w}_launches.csv > gpurun_out/r2_${w}_launch_shares.txt
done
full() {  # name, kernel regex, count, bench args...
  local name=$1 rx=$2 cnt=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -c $cnt -f -o /tmp/$name $B "$@" > gpurun_out/${name}_ncu.log 2>&1; echo ncu full $name rc=$?
  python tools/ncu_summary.py raw /tmp/$name.ncu-rep > gpurun_out/${name}_ncu_summary.txt
  ncu -i /tmp/$name.ncu-rep --page details --csv 2>/dev/null | gzip > gpurun_out/${name}_details.csv.gz
  rm -f /tmp/$name.ncu-rep
}
full r2_alarm_kcount k_count 3 --steps 1 --warmup 1
full r2_diabetes_kcount k_count 4 --workload diabetes --steps 1 --warmup 1
full r2_diabetes_kderive k_derive 3 --workload diabetes --steps 1 --warmup 1
full r2_pigs_kcount k_count 1 --workload pigs --steps 1 --warmup 1
du -sh gpurun_out; ls -la gpurun_out
