#!/bin/bash
# r02 call U: the full default bench line at N=1 (all batched configs) at HEAD, exactly as the driver runs it
set -x
mkdir -p gpurun_out
timeout 560 python bench.py --gpus 1 > gpurun_out/u_bench_n1_full.json 2> gpurun_out/u_bench_n1_full.err
echo "rc=$?"
tail -3 gpurun_out/u_bench_n1_full.err
python - <<'PY'
import json
for l in open("gpurun_out/u_bench_n1_full.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d.get(k) for k in ("value", "ms_per_step", "batch1_transport", "parity_ok", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
        print(d.get("concurrent_callers")); print(d.get("session"))
        for c, v in (d.get("configs") or {}).items():
            if "error" in v or "skipped" in v: print(c, v); continue
            for r in v.get("runs", []):
                rf = r.get("roofline", {})
                print("   ", c, r.get("k"), round(r.get("ms_per_batch"), 3), round(r.get("value")), r.get("parity_ok"), r.get("tc_fallbacks"), rf.get("frac"), rf.get("hbm", {}).get("frac"), r.get("recall_at_10"))
PY
