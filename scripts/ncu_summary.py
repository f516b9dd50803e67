#!/usr/bin/env python
"""Condense an .ncu-rep (one kernel, --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python scripts/ncu_summary.py gpurun_out/prof_tb.ncu-rep profiles/r01_tb_filter_kernel"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.avg.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall = [n for n in h if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
res = []
for r in rows[2:]:
    d = {}
    for n in want + stall:
        if n in h:
            i = h.index(n)
            d[n] = (r[i] + (" " + units[i] if units[i] else "")).strip()
    res.append(d)
json.dump(res, open(out + "_summary.json", "w"), indent=1)
with open(out + "_summary.md", "w") as f:
    for k, d in enumerate(res):
        f.write("## launch %d\n\n| metric | value |\n|---|---|\n" % k)
        for n, v in d.items():
            f.write("| %s | %s |\n" % (n, v))
        f.write("\n")
print(json.dumps(res[-1], indent=1))
