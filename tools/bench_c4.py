#!/usr/bin/env python
"""BASELINE config 4: 50M x 1024 f16, row-sharded across the ranks, batch-64 queries, top-k, tensor-core path,
per-shard exact top-k exchanged once per batch.  Run plain for 1 GPU or under torch.distributed.run for N GPUs.
Prints one JSON line: batches/s, queries/s, ms/batch (device events, max over ranks), fallbacks."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=64)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()
    import torch, torch.distributed as dist
    cg = ge.load_package()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        buf = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0: buf.copy_(torch.frombuffer(bytearray(cg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0); uid = bytes(buf.cpu().numpy().tobytes())
    b, e = cg.shard_range(a.rows, world, rank)
    ix = cg.Index(a.dim, cg.F16, device=lr, rank=rank, world=world, nccl_unique_id=uid, row_offset=b) if world > 1 else cg.Index(a.dim, cg.F16, device=lr)
    ix.reserve(e - b); ix.fill_synthetic(e - b, 0xC0DE6A9F, True)
    for o in a.opt:
        kk, v = o.split("="); ix.set_option(kk, int(v))
    qs = torch.from_numpy(np.random.default_rng(1).standard_normal((a.iters + 2, a.nq, a.dim)).astype(np.float32)).to(dev)
    o_r = torch.empty((a.nq, a.k), dtype=torch.int64, device=dev); o_s = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    o_c = torch.empty((a.nq,), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    def run(i): ix.search_device(qs[i].data_ptr(), a.nq, a.k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), cg.COSINE, st.cuda_stream, cg.PATH_TENSOR)
    def barrier():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    for i in range(2): run(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(a.iters): run(2 + i)
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1) / a.iters
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    s = ix.stats()
    if rank == 0:
        print(json.dumps({"config": f"C4: {a.rows} x {a.dim} f16, batch-{a.nq}, top-{a.k}, row-sharded x{world}", "n_gpus": world, "ms_per_batch": round(ms, 3),
                          "queries_per_s": round(a.nq / ms * 1e3, 1), "GBps_per_gpu": round((e - b) * a.dim * 2 / ms / 1e6, 1),
                          "tc_batches": int(s.tc_batches), "tc_fallbacks": int(s.tc_fallbacks), "first_hit": [int(o_r[0, 0]), float(o_s[0, 0])], "opts": a.opt}), flush=True)
    ix.close()
    if world > 1: dist.destroy_process_group()

if __name__ == "__main__":
    main()
