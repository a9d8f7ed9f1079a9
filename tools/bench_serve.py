#!/usr/bin/env python
"""Resident session vs launch-per-query on one GPU: end-to-end (host buffers) and pipelined device-resident queries per
second for the C2 shape and for the per-GPU shards C2 has at 2/4/8 GPUs (rows / N)."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=768); ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--rows", type=int, nargs="*", default=[1_000_000, 500_000, 250_000, 125_000])
    ap.add_argument("--chunks", type=int, nargs="*", default=[0])
    args = ap.parse_args()
    cg = ge.load_package()
    import torch
    rng = np.random.default_rng(2)
    for rows in args.rows:
        ix = cg.Index(args.dim)
        ix.reserve(rows); ix.fill_synthetic(rows, 0xC0DE6A9F, True)
        ix.set_option("coalesce", 0)
        qs = rng.standard_normal((64, args.dim)).astype(np.float32)
        dq = torch.from_numpy(qs).cuda()
        o_r = torch.empty((64, args.k), dtype=torch.int64, device="cuda"); o_s = torch.empty((64, args.k), dtype=torch.float32, device="cuda")
        o_c = torch.empty((64,), dtype=torch.int32, device="cuda")
        alg = rows * args.dim * 4 + rows * 4
        rec = {"rows": rows, "dim": args.dim, "k": args.k, "ideal_us_at_6555GBs": round(alg / 6555e3, 1)}
        # launch per query: end to end (host buffers) and device-resident back to back
        bufs = ix.make_search_buffers(1, args.k)
        for i in range(20): ix.search_into(qs[i % 64][None, :], bufs)
        t0 = time.perf_counter()
        for i in range(args.steps): ix.search_into(qs[i % 64][None, :], bufs)
        rec["launch_e2e_us"] = round((time.perf_counter() - t0) / args.steps * 1e6, 2)
        st = torch.cuda.Stream()
        for i in range(20): ix.search_device(dq[i % 64].data_ptr(), 1, args.k, o_r[0].data_ptr(), o_s[0].data_ptr(), o_c[0].data_ptr(), cg.COSINE, st.cuda_stream, cg.PATH_EXACT)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(args.steps): ix.search_device(dq[i % 64].data_ptr(), 1, args.k, o_r[i % 64].data_ptr(), o_s[i % 64].data_ptr(), o_c[i % 64].data_ptr(), cg.COSINE, st.cuda_stream, cg.PATH_EXACT)
        e1.record(st); torch.cuda.synchronize()
        rec["launch_device_us"] = round(e0.elapsed_time(e1) / args.steps * 1e3, 2)
        for chunk in args.chunks:
            s = cg.ServeSession(ix, args.k)
            if chunk: s.set("contig", chunk - 1)            # --chunks 1 2 -> interleaved / contiguous tile runs per CTA (A/B knob)
            qrows = [np.ascontiguousarray(qs[i]) for i in range(64)]
            for i in range(20): s.search_raw(qrows[i % 64])
            t0 = time.perf_counter()
            for i in range(args.steps): s.search_raw(qrows[i % 64])
            e2e = (time.perf_counter() - t0) / args.steps * 1e6
            # pipelined device-resident submissions (the session's own flow control keeps <= 6 in flight)
            for i in range(20): t = s.submit_device(dq[i % 64].data_ptr(), o_r[i % 64].data_ptr(), o_s[i % 64].data_ptr(), o_c[i % 64].data_ptr())
            s.wait(t)
            t0 = time.perf_counter()
            for i in range(args.steps): t = s.submit_device(dq[i % 64].data_ptr(), o_r[i % 64].data_ptr(), o_s[i % 64].data_ptr(), o_c[i % 64].data_ptr())
            s.wait(t)
            dev = (time.perf_counter() - t0) / args.steps * 1e6
            rec[f"serve_e2e_us_chunk{chunk}"] = round(e2e, 2); rec[f"serve_device_us_chunk{chunk}"] = round(dev, 2)
            rec["serve_launches"] = s.stats()["launches"]
            s.close()
        # the launch path once more, AFTER the sessions: separates a real difference from the board's power state drifting
        e0.record(st)
        for i in range(args.steps): ix.search_device(dq[i % 64].data_ptr(), 1, args.k, o_r[i % 64].data_ptr(), o_s[i % 64].data_ptr(), o_c[i % 64].data_ptr(), cg.COSINE, st.cuda_stream, cg.PATH_EXACT)
        e1.record(st); torch.cuda.synchronize()
        rec["launch_device_us_again"] = round(e0.elapsed_time(e1) / args.steps * 1e3, 2)
        print(json.dumps(rec), flush=True)
        ix.close()

if __name__ == "__main__":
    main()
