#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (ncu --set full): time, DRAM bytes, issue/tensor utilisation and the warp-stall table.
  python tools/ncu_stalls.py gpurun_out/x.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ik = h.index("Kernel Name")
def col(name):
    return h.index(name) if name in h else None
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]
stall = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if flt and flt not in r[ik]:
        continue
    print("==", r[ik][:90])
    for k in keep:
        i = col(k)
        if i is not None:
            print("   %-70s %s %s" % (k, r[i], units[i]))
    st = sorted(((float(r[h.index(c)].replace(",", "")), c.split("stalled_")[1].split("_per_issue")[0]) for c in stall), reverse=True)
    print("   stalls per issue: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:8]))
