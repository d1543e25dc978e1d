#!/usr/bin/env python
"""Summaries of ncu captures for profiles/.

  python tools/ncu_summary.py raw  <file.ncu-rep> [kernel-regex]   -> metric table per captured launch
  python tools/ncu_summary.py shares <launches.csv>                 -> per-kernel share of device time

`raw` runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps the metrics DESIGN.md
quotes; `shares` reads the CSV of `ncu --metrics gpu__time_duration.sum --csv --log-file ...`.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__cycles_elapsed.avg",
]


def raw(rep, pattern=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    if pattern:
        data = [r for r in data if re.search(pattern, r[col["Kernel Name"]])]
    lines = ["Kernel Name | " + " | ".join(r[col["Kernel Name"]] for r in data),
             "Grid Size | " + " | ".join(r[col["Grid Size"]] for r in data),
             "Block Size | " + " | ".join(r[col["Block Size"]] for r in data)]
    for m in KEEP:
        if m in col:
            lines.append(f"{m:100s} {units[col[m]]:>16s} | " + " | ".join(r[col[m]] for r in data))
    return "\n".join(lines)


def shares(path):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("bic::", "")
        a = agg.setdefault(name, [0.0, 0, 0.0])
        a[0] += us
        a[1] += 1
        a[2] = max(a[2], us)
    total = sum(a[0] for a in agg.values())
    lines = []
    for name, (us, cnt, mx) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{us:12.1f} us {cnt:5d} {100 * us / total:5.1f}%  max {mx:9.1f}  {name}")
    lines.append(f"total us {total:.3f} launches {sum(a[1] for a in agg.values())}")
    return "\n".join(lines)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "raw":
        print(raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None))
    elif len(sys.argv) == 3 and sys.argv[1] == "shares":
        print(shares(sys.argv[2]))
    else:
        sys.exit(__doc__)
