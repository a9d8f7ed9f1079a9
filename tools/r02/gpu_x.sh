#!/bin/bash
# r02 call X: ncu launch list (gpu__time_duration.sum, no clock control) of the bench command's C2 path — kernel share of the step
set -x
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/x_c2_launches.csv python bench.py --steps 20 --warmup 3 --configs none --no-extras --no-cpu-baseline > gpurun_out/x_bench_under_ncu.log 2>&1
echo "rc=$?"
tail -2 gpurun_out/x_bench_under_ncu.log | cut -c1-200
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/x_c2_launches.csv")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; idx = {x: i for i, x in enumerate(h)}
agg = collections.defaultdict(list)
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[idx["Metric Name"]] != "gpu__time_duration.sum": continue
    agg[r[idx["Kernel Name"]][:60]].append(float(r[idx["Metric Value"]].replace(",", "")))
for k, v in agg.items():
    print(f"{k:62s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.2f} us  last={v[-1]/1e3:9.2f} us")
PY
