#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu --page source --csv`.  Usage: ncu_hot.py rep [top=40]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = body[i]
    n = int(r[idx["# Samples"]] or 0)
    st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {n:7d} {100*n/tot:5.1f}%  ex={r[idx['Instructions Executed']]:>9}  {r[idx['Source']].strip()[:90]:90s} {' '.join(f'{c}:{v}' for v, c in st if v)}")
