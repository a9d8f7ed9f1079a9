#!/usr/bin/env python
"""Exact-order kernel vs tensor path for small batches through the host-buffer API: ms per call for nq in {1,2,4,8,16,32}
at several index sizes.  Decides where AUTO should switch (tc_min_batch)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=768); ap.add_argument("--k", type=int, default=10); ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    cg = ge.load_package()
    rng = np.random.default_rng(3)
    for dtype_name, dt in (("f32", cg.F32), ("f16", cg.F16)):
        for rows in (65_536, 262_144, 1_000_000, 4_000_000):
            ix = cg.Index(args.dim, dt)
            ix.reserve(rows); ix.fill_synthetic(rows, 0xC0DE6A9F, True)
            ix.set_option("coalesce", 0)
            for nq in (1, 2, 4, 8, 16, 32):
                qs = rng.standard_normal((nq, args.dim)).astype(np.float32)
                rec = {"dtype": dtype_name, "rows": rows, "nq": nq}
                for name, path in (("exact", cg.PATH_EXACT), ("tensor", cg.PATH_TENSOR)):
                    try:
                        for _ in range(3): ix.search(qs, args.k, cg.COSINE, path=path)
                        f0 = ix.stats().tc_fallbacks
                        t0 = time.perf_counter()
                        for _ in range(args.reps): ix.search(qs, args.k, cg.COSINE, path=path)
                        rec[name + "_ms"] = round((time.perf_counter() - t0) / args.reps * 1e3, 4)
                        if name == "tensor": rec["fallbacks"] = ix.stats().tc_fallbacks - f0
                    except cg.CgvecError as e:
                        rec[name + "_ms"] = None; rec[name + "_err"] = str(e)[:60]
                print(json.dumps(rec), flush=True)
            ix.close()

if __name__ == "__main__":
    main()
