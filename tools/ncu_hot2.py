#!/usr/bin/env python
"""Top SASS instructions by stall samples for every kernel section of `ncu --page source --csv` (multi-launch reports).
Usage: ncu_hot2.py rep [top=40] [section-index]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = int(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = []
for r in rows:
    if r and r[0] == "Kernel Name": secs.append({"name": r[1], "hdr": None, "body": []})
    elif secs and secs[-1]["hdr"] is None: secs[-1]["hdr"] = r
    elif secs: secs[-1]["body"].append(r)
for si, sec in enumerate(secs):
    hdr, body = sec["hdr"], sec["body"]
    idx = {h: i for i, h in enumerate(hdr)}
    if "# Samples" not in idx: continue
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
    print(f"== section {si}: {sec['name'][:60]} rows={len(body)} total samples {tot}")
    if want is not None and si != want: continue
    order = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        n = int(r[idx["# Samples"]] or 0)
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {n:7d} {100*n/max(tot,1):5.1f}%  ex={r[idx['Instructions Executed']]:>9}  {r[idx['Source']].strip()[:80]:80s} {' '.join(f'{c}:{v}' for v, c in st if v)}")
