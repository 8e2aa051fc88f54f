#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep <substring of kernel name> [inst|samples] [n]"""
import csv
import subprocess
import sys
rep, pat = sys.argv[1], sys.argv[2]
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
out, fn, take = [], None, False
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "File Path":
        fn = r[1].split("/")[-1]
    elif r and r[0] == "Function Name":
        take = pat in r[1]
    elif take and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            out.append((int(r[7]), int(r[6]), fn, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
ti, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print("instructions", ti, "samples", ts)
for o in sorted(out, key=lambda o: -(o[1] if key == "samples" else o[0]))[:top]:
    print("%5.1f%% inst  %5.1f%% samples  %s:%s  %s" % (100 * o[0] / ti, 100 * o[1] / ts, o[2], o[3], o[4]))
