#!/usr/bin/env python
"""int8 quantised scan (cgvec_search_i8 = search_optimized, optimization.rs:63-150) throughput on one GPU:
rows x dim codes (1 byte per element), batch-1 query, top-limit.  Algorithmic bytes per query = rows*dim + 4*rows."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--iters", type=int, default=200)
    a = ap.parse_args()
    cg = ge.load_package()
    ix = cg.Index(a.dim, cg.F32)
    ix.reserve(a.rows); ix.fill_synthetic(a.rows, 0xC0DE6A9F, True)
    ix.quantize_i8()
    rng = np.random.default_rng(0)
    qs = (rng.standard_normal((a.iters + 5, a.dim)) / np.sqrt(a.dim)).astype(np.float32)
    for i in range(5): ix.search_optimized(qs[i], a.limit)
    t0 = time.perf_counter()
    for i in range(a.iters): ix.search_optimized(qs[5 + i], a.limit)
    dt = (time.perf_counter() - t0) / a.iters
    alg = a.rows * a.dim + 4 * a.rows
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"rows": a.rows, "dim": a.dim, "limit": a.limit, "ms_per_query_e2e": round(dt * 1e3, 4), "qps_e2e": round(1 / dt, 1),
                      "algorithmic_bytes": alg, "GBps_on_e2e_time": round(alg / dt / 1e9, 1), "frac_of_measured_hbm_on_e2e_time": round(alg / dt / 1e9 / peak, 3)}))
    ix.close()

if __name__ == "__main__":
    main()
