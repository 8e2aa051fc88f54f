"""Summarise an `ncu --page source --csv --print-source sass` export: instruction and stall-sample shares by region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; rows = rows[2:]
ia, isrc, iinst, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[iinst]) for r in rows); tot_s = sum(int(r[ismp]) for r in rows)
print("total inst", tot_i, "samples", tot_s)
# overall stall reasons
agg = {hdr[i]: sum(int(r[i] or 0) for r in rows) for i in stall}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
mode = sys.argv[2] if len(sys.argv) > 2 else "top"
if mode == "top":
    for k, r in sorted(enumerate(rows), key=lambda kr: -int(kr[1][ismp]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        top = sorted(((hdr[i], int(r[i] or 0)) for i in stall), key=lambda kv: -kv[1])[:2]
        print(k, r[isrc][:70], "inst", r[iinst], "smp", r[ismp], top)
elif mode == "dump":
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for k in range(lo, hi):
        r = rows[k]
        print(k, r[isrc][:80], r[iinst], r[ismp])
elif mode == "blocks":   # cumulative instruction share in blocks of N sass lines
    n = int(sys.argv[3])
    for b in range(0, len(rows), n):
        bi = sum(int(r[iinst]) for r in rows[b:b+n]); bs = sum(int(r[ismp]) for r in rows[b:b+n])
        print(b, f"inst {100*bi/tot_i:5.1f}%  samples {100*bs/tot_s:5.1f}%")
