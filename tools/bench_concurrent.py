#!/usr/bin/env python
"""Concurrent batch-1 callers against one index (multi_vector_search's shape, search.rs:347-361): queries/s for T caller threads
with the group commit of cgvec_search_ex switched on and off.  C2 matrix by default (1M x 768 f32, top-10)."""
import argparse, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def run(cg, ix, qs, k, threads, reps):
    bufs = [ix.make_search_buffers(1, k) for _ in range(threads)]
    start = threading.Barrier(threads + 1)
    def work(i):
        start.wait()
        for r in range(reps):
            ix.search_into(qs[(i * reps + r) % len(qs)][None, :], bufs[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    start.wait(); t0 = time.perf_counter()
    [t.join() for t in th]
    return threads * reps / (time.perf_counter() - t0)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000); ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10); ap.add_argument("--reps", type=int, default=200)
    ap.add_argument("--dtype", default="f32")
    args = ap.parse_args()
    cg = ge.load_package()
    ix = cg.Index(args.dim, cg.F32 if args.dtype == "f32" else cg.F16)
    ix.reserve(args.rows); ix.fill_synthetic(args.rows, 0xC0DE6A9F, True)
    qs = np.random.default_rng(1).standard_normal((256, args.dim)).astype(np.float32)
    out = {"rows": args.rows, "dim": args.dim, "k": args.k, "dtype": args.dtype, "runs": []}
    for co, cmax in ((0, 16), (1, 4), (1, 16)):
        ix.set_option("coalesce", co); ix.set_option("coalesce_max", cmax)
        for T in (1, 4, 16, 32):
            run(cg, ix, qs, args.k, T, 20)
            s0 = ix.stats()
            qps = run(cg, ix, qs, args.k, T, args.reps)
            s1 = ix.stats()
            out["runs"].append({"coalesce": co, "coalesce_max": cmax, "threads": T, "qps": round(qps, 1),
                                "coalesced_batches": s1.coalesced_batches - s0.coalesced_batches,
                                "coalesced_queries": s1.coalesced_queries - s0.coalesced_queries,
                                "tc_batches": s1.tc_batches - s0.tc_batches, "launches": s1.kernel_launches - s0.kernel_launches})
            print(json.dumps(out["runs"][-1]), flush=True)
    ix.close()

if __name__ == "__main__":
    main()
