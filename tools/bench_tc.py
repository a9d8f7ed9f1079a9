#!/usr/bin/env python
"""Batched (tensor-core path) throughput on one GPU: rows x dim f16, nq-query batches, top-k.
Reports ms/batch, qps, achieved HBM GB/s (rows*dim*2 bytes per batch pass) and the exact-kernel fallback count."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=6_250_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=64)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--path", default="tensor")
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--opt", action="append", default=[])
    args = ap.parse_args()
    cg = ge.load_package()
    if os.environ.get("CGVEC_AB_LIB"):                      # A/B runs: load another build of the library (tools/r02/gpu_k.sh)
        cg._build.LIB = os.environ["CGVEC_AB_LIB"]
        cg.load_library(build=False)
    import torch
    ix = cg.Index(args.dim, cg.F16 if args.dtype == "f16" else cg.F32)
    ix.reserve(args.rows)
    ix.fill_synthetic(args.rows, 0xC0DE6A9F, True)
    for o in args.opt:
        kk, v = o.split("="); ix.set_option(kk, int(v))
    rng = np.random.default_rng(0)
    qs = torch.from_numpy(rng.standard_normal((args.iters + 2, args.nq, args.dim)).astype(np.float32)).cuda()
    o_r = torch.empty((args.nq, args.k), dtype=torch.int64, device="cuda")
    o_s = torch.empty((args.nq, args.k), dtype=torch.float32, device="cuda")
    o_c = torch.empty((args.nq,), dtype=torch.int32, device="cuda")
    path = {"tensor": cg.PATH_TENSOR, "exact": cg.PATH_EXACT, "auto": cg.PATH_AUTO}[args.path]
    st = torch.cuda.Stream()
    for i in range(2):
        ix.search_device(qs[i].data_ptr(), args.nq, args.k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), cg.COSINE, st.cuda_stream, path)
    torch.cuda.synchronize()
    l0 = ix.stats().kernel_launches
    ix.set_option("reset_timing", 1); ix.set_option("timing", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    for i in range(args.iters):
        ix.search_device(qs[2 + i].data_ptr(), args.nq, args.k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), cg.COSINE, st.cuda_stream, path)
    e1.record(st)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    s = ix.stats()
    ms = e0.elapsed_time(e1) / args.iters
    alg = args.rows * args.dim * (2 if args.dtype == "f16" else 4)
    passes = -(-args.nq // 128) if args.path != "exact" else args.nq
    print(json.dumps({"rows": args.rows, "dim": args.dim, "nq": args.nq, "k": args.k, "path": args.path, "dtype": args.dtype, "ms_per_batch": round(ms, 3),
                      "wall_ms_per_batch": round(wall / args.iters * 1e3, 3), "qps": round(args.nq / ms * 1e3, 1),
                      "GBps_per_pass": round(alg * passes / ms / 1e6, 1), "scan_ms_events": round(s.scan_ms_total / max(s.scans_timed, 1), 3),
                      "launches_per_batch": (s.kernel_launches - l0) / args.iters, "tc_batches": s.tc_batches, "tc_fallbacks": s.tc_fallbacks,
                      "opts": args.opt}), flush=True)
    ix.close()

if __name__ == "__main__":
    main()
