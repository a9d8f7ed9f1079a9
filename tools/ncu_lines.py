#!/usr/bin/env python
"""Stall samples per CUDA source line from an .ncu-rep (cuda,sass view).  Usage: ncu_lines.py rep [top=40] [launch=last]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
launches = []          # each: dict (file,line)->(samples, src, stalls)
cur = None; fname = None; hdr = None; seen_files = set()
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if cur is None or fname in seen_files: cur = {}; launches.append(cur); seen_files = set()
        seen_files.add(fname); continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or cur is None: continue
    if r[0] not in ("", "-") and r[0].isdigit():
        si = hdr.index("# Samples")
        try: n = int(r[si] or 0)
        except ValueError: n = 0
        stall = {h[6:]: int(r[i] or 0) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and i < len(r) and r[i].isdigit()}
        cur[(fname, int(r[0]))] = (n, r[1].strip(), stall)
sel = launches[-1] if len(sys.argv) <= 3 else launches[int(sys.argv[3])]
print("launch sections:", len(launches))
tot = sum(v[0] for v in sel.values())
print("total samples", tot)
for (f, ln), (n, src, st) in sorted(sel.items(), key=lambda kv: -kv[1][0])[:top]:
    s3 = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{f}:{ln:<4d} {n:7d} {100*n/max(tot,1):5.1f}%  {src[:95]:95s} {s3}")
