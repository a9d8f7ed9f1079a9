#!/usr/bin/env python
"""Times the exact-order scan kernel (device events inside the library) across launch geometries.
Usage: python tools/sweep.py [--rows 1000000 --dim 768 --dtype f32 --k 10 --iters 30]"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=1)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--tiles", default="8,16,32")
    ap.add_argument("--stages", default="2,3,4,6,8")
    ap.add_argument("--hints", default="0,1")
    ap.add_argument("--syncs", default="8")
    ap.add_argument("--metric", default="cosine")
    args = ap.parse_args()
    cg = ge.load_package()
    import torch
    dt = cg.F32 if args.dtype == "f32" else cg.F16
    es = 4 if args.dtype == "f32" else 2
    ix = cg.Index(args.dim, dt)
    ix.reserve(args.rows)
    ix.fill_synthetic(args.rows, 0xC0DE6A9F, True)
    rng = np.random.default_rng(0)
    qs = torch.from_numpy(rng.standard_normal((64, args.nq, args.dim)).astype(np.float32)).cuda()
    o_r = torch.empty((args.nq, args.k), dtype=torch.int64, device="cuda")
    o_s = torch.empty((args.nq, args.k), dtype=torch.float32, device="cuda")
    o_c = torch.empty((args.nq,), dtype=torch.int32, device="cuda")
    metric = {"cosine": cg.COSINE, "dot": cg.DOT, "l2": cg.L2}[args.metric]
    alg = args.rows * args.dim * es + (args.rows * 4 if args.metric == "cosine" else 0)
    ix.set_option("max_batch", args.nq)
    results = []
    for tile, stages, hint, sync in itertools.product([int(x) for x in args.tiles.split(",")], [int(x) for x in args.stages.split(",")],
                                                     [int(x) for x in args.hints.split(",")], [int(x) for x in args.syncs.split(",")]):
        ix.set_option("tile_rows", tile); ix.set_option("stages", stages); ix.set_option("l2_hint", hint); ix.set_option("sync_interval", sync)
        try:
            for i in range(3):
                ix.search_device(qs[i].data_ptr(), args.nq, args.k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), metric)
            torch.cuda.synchronize()
            ix.set_option("reset_timing", 1); ix.set_option("timing", 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for i in range(args.iters):
                ix.search_device(qs[(3 + i) % 64].data_ptr(), args.nq, args.k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), metric)
            torch.cuda.synchronize()
            st = ix.stats()
            ix.set_option("timing", 0)
            ms = st.scan_ms_total / max(st.scans_timed, 1)
            res = {"tile": tile, "stages_req": stages, "stages": st.stages, "tile_rows": st.tile_rows, "hint": hint, "sync": sync, "smem": st.smem_bytes,
                   "scan_ms": round(ms, 4), "GBps": round(alg / ms / 1e6, 1)}
        except cg.CgvecError as e:
            res = {"tile": tile, "stages_req": stages, "hint": hint, "error": e.msg}
        results.append(res)
        print(json.dumps(res), flush=True)
    ok = [r for r in results if "GBps" in r]
    if ok:
        best = max(ok, key=lambda r: r["GBps"])
        print("BEST", json.dumps(best), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
