#!/usr/bin/env python
"""Per-step device timeline (option "trace") of back-to-back batch-1 searches: scan duration, merge/exchange duration and
the gaps between consecutive kernels, for a given shard size.  Usage: python tools/trace_steps.py [--rows 125000] [--steps 200]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=125_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    cg = ge.load_package()
    import torch
    ix = cg.Index(a.dim)
    ix.fill_synthetic(a.rows, 0xC0DE6A9F, True)
    for o in a.opt:
        k, v = o.split("="); ix.set_option(k, int(v))
    qs = torch.from_numpy(np.random.default_rng(0).standard_normal((a.steps + 20, a.dim)).astype(np.float32)).cuda()
    o_r = torch.empty((a.steps + 20, 10), dtype=torch.int64, device="cuda"); o_s = torch.empty((a.steps + 20, 10), dtype=torch.float32, device="cuda")
    o_c = torch.empty((a.steps + 20,), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    def run(i): ix.search_device(qs[i].data_ptr(), 1, 10, o_r[i].data_ptr(), o_s[i].data_ptr(), o_c[i].data_ptr(), cg.COSINE, st.cuda_stream)
    for i in range(20): run(i)
    torch.cuda.synchronize()
    ix.set_option("trace", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(20, 20 + a.steps): run(i)
    e1.record(st)
    torch.cuda.synchronize()
    tr = ix.trace()
    scans = tr[tr[:, 0] == 1]; others = tr[tr[:, 0] != 1]
    dur_scan = (scans[:, 2] - scans[:, 1]) / 1e3
    dur_other = (others[:, 2] - others[:, 1]) / 1e3
    period = np.diff(scans[:, 1]) / 1e3
    gap_after_scan = (others[: len(scans), 1] - scans[: len(others), 2]) / 1e3
    gap_scan_to_scan = (scans[1:, 1] - scans[:-1, 2]) / 1e3
    f = lambda x: f"median {np.median(x):.2f} p10 {np.percentile(x,10):.2f} p90 {np.percentile(x,90):.2f}"
    print(f"rows={a.rows} opts={a.opt} step(event)={e0.elapsed_time(e1)/a.steps*1e3:.2f} us")
    print(" scan kernel (first CTA start -> last CTA end) us:", f(dur_scan))
    print(" merge/exchange kernel us:                        ", f(dur_other))
    print(" scan start -> next scan start us:                ", f(period))
    print(" scan end -> merge start us:                      ", f(gap_after_scan))
    print(" scan end -> next scan start us:                  ", f(gap_scan_to_scan))
    ix.close()

if __name__ == "__main__":
    main()
