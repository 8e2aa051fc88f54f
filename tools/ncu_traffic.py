#!/usr/bin/env python
"""Per-launch DRAM traffic of one forward + fused decode/NMS step of the bench workload.

  on the GPU box:   python tools/ncu_traffic.py capture gpurun_out/traffic.csv
  here:             python tools/ncu_traffic.py summarise gpurun_out/traffic.csv profiles/r2_traffic.json

`capture` runs tools/prof_fwd.py under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--clock-control none` and keeps the launches of the LAST step; `summarise` names them after the plan's launch groups (they
run in plan order, then decode+NMS) and writes {"per_launch_bytes": {unit: read+write}, "launches": [...]}, which bench.py
reads for `roofline.traffic` and the per-stage `traffic_ratio`."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS = 3
# the forward / post-processing kernels only (weight packing and torch fills also launch during the run)
KERNELS = "regex:^(stem|s1c_|s2c_|pw3_|tail_|tc_|decode_nms|nms_|decode_)"


def group_labels():
    """Launch-group labels in plan order (same rule as bench.py), from the stage names / groups of a CPU-side plan."""
    sys.path.insert(0, ROOT)
    import ctypes
    import yfv2  # noqa: F401
    import yfv2_engine as eng
    L = eng.lib()
    h = ctypes.c_void_p()
    assert L.yfv2_plan_create(ctypes.byref(h), 0, 256, 352, 352, 3, 80, 0) == 0
    names, groups = [], []
    while True:
        nm = L.yfv2_plan_stage_name(h, len(names))
        if nm is None:
            break
        groups.append(L.yfv2_plan_stage_group(h, len(names)))
        names.append(nm.decode())
    L.yfv2_plan_destroy(h)
    out = []
    for nm, g in zip(names, groups):
        if out and out[-1][0] == g:
            out[-1][1].append(nm)
        else:
            out.append([g, [nm]])
    labels = []
    for _, ns in out:
        units = []
        for n in ns:
            if n.split("/")[0] not in units:
                units.append(n.split("/")[0])
        lab = units[0] if len(units) == 1 else "%s-%s" % (units[0], units[-1].split(".")[-1])
        labels.append(lab if len(ns) == 1 or len(units) > 1 else ns[0])
    # K=96 stride-2 block: two launches of one unit keep their own names ("stage4.0/pw1", "stage4.0/dwpw")
    flat = []
    for (_, ns), lab in zip(out, labels):
        flat.append(lab if "/" not in ns[0] or len(ns) > 1 else ns[0])
    return flat


def capture(path):
    labels = group_labels()
    per_step = len(labels) + 1                       # + fused decode/NMS
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", KERNELS, "-s", str((STEPS - 1) * per_step), "-c", str(per_step), "--csv", "--log-file", path,
           sys.executable, os.path.join(ROOT, "tools", "prof_fwd.py"), str(STEPS)]
    print(" ".join(cmd), flush=True)
    sys.exit(subprocess.call(cmd))


def summarise(path, out):
    labels = group_labels() + ["decode+nms"]
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    hdr = next(r for r in csv.reader(open(path)) if r and r[0] == "ID")
    iid, ik, im, iv, iu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik].split("(")[0][-60:]})
        d[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
    ids = sorted(launches)
    assert len(ids) == len(labels), "captured %d launches, the plan has %d" % (len(ids), len(labels))
    res, per = [], {}
    for lab, i in zip(labels, ids):
        d = launches[i]
        b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        res.append({"launch": lab, "kernel": d["kernel"], "us_under_ncu": round(d["gpu__time_duration.sum"], 2),
                    "dram_read_MB": round(d["dram__bytes_read.sum"] / 1e6, 2), "dram_write_MB": round(d["dram__bytes_write.sum"] / 1e6, 2)})
        unit = lab.split("/")[0]
        per[unit] = per.get(unit, 0.0) + b
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch of one bench step (batch 256 @352x352), ncu --clock-control none; "
                       "regenerate with tools/ncu_traffic.py",
               "per_launch_bytes": {k: int(v) for k, v in per.items()}, "launches": res}, open(out, "w"), indent=1)
    for r in res:
        print("%-16s %-44s %8.1f us  read %8.1f MB  write %8.1f MB" % (r["launch"], r["kernel"][-44:], r["us_under_ncu"], r["dram_read_MB"], r["dram_write_MB"]))


if __name__ == "__main__":
    if sys.argv[1] == "capture":
        capture(sys.argv[2])
    else:
        summarise(sys.argv[2], sys.argv[3])
