#!/usr/bin/env python
"""bench.py — queries/sec and achieved HBM GB/s (or tensor TFLOP/s) of the embedding similarity-search hot path.

Headline workload (BASELINE.json configs[1], "C2"): 1M x 768 f32 resident matrix, batch-1 query, top-10 cosine,
exact-order scan.  A "step" is one query: one pass of the scan + top-k over the whole index.  At N > 1 GPUs the SAME
index is row-sharded across the ranks (strong scaling; one process per GPU), each rank scans its shard and the
per-shard top-k lists are exchanged once per query, so `value` stays "queries/sec over the 1M x 768 index".

The same JSON line carries a `configs` object with one sub-record per remaining GPU config of BASELINE.json, each run
on the same N ranks (row-sharded), each with its own roofline and an oracle parity check:
  c3   10M x 768 f16, batch-256, top-100, tcgen05 path
  c4   50M x 1024 f16, batch-64, top-100 and top-10, tcgen05 path (the north_star's 8-GPU target shape)
  c5   100M x 384 f32 (TF32 tensor path), streaming 1024-query batches from host memory, recall@10 vs the oracle

  python bench.py [--gpus N --steps K --warmup W] [--configs c3,c4,c5 | --configs none]
  python bench.py --impl reference [...]        the reference's CPU path (oracle port of
                                                ParallelVectorOps::parallel_top_k_search, all host threads)

Prints ONE JSON line (rank 0).  See the task contract for field meanings.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, dim, dtype, batch, k)
    "c1": (10_000, 768, "f32", 1, 10),
    "c2": (1_000_000, 768, "f32", 1, 10),
}
BATCHED = {
    # name: rows, dim, dtype, batch, ks, streaming
    "c3": dict(rows=10_000_000, dim=768, dtype="f16", batch=256, ks=(100,), streaming=False,
               what="10M x 768 f16, batch-256 queries, top-100 cosine, tcgen05 path (BASELINE configs[2])"),
    "c4": dict(rows=50_000_000, dim=1024, dtype="f16", batch=64, ks=(100, 10), streaming=False,
               what="50M x 1024 f16 row-sharded, batch-64 queries, top-k cosine, tcgen05 path, one exchange per batch (BASELINE configs[3])"),
    "c5": dict(rows=100_000_000, dim=384, dtype="f32", batch=1024, ks=(10,), streaming=True,
               what="100M x 384 f32 row-sharded, streaming 1024-query batches from host memory, top-10 cosine, TF32 tcgen05 path (BASELINE configs[4])"),
}
SEED_ROWS, SEED_QUERIES = 0xC0DE6A9F, 0x5EED0001
METRIC = "queries/sec and HBM GB/s over N x d embeddings (1M x 768 f32, batch-1, top-10 cosine) vs CPU ref"
FULL_ORACLE_MAX_LOCAL_ROWS = 13_000_000     # above this per-rank shard size the full CPU oracle pass no longer fits a bench run


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)),
                "source": "measured (MEASURED_PEAKS.json: copy read+write GB/s; cuBLAS bf16 burst / sustained TFLOP/s)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s burst, ~1.4 sustained)"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed regions run."""

    def __init__(self, device_index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.period = period_s
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.stop_flag = threading.Event()
        self.ok = False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self, t0: float, t1: float):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        mhz = sorted(s[1] for s in inside)
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside)}


def agreed_count(max_over_ranks, local_seconds_per_unit, target_s, lo=1, hi=500):
    """How many units (pre-warm rounds, timed batches) fit `target_s`, identical on every rank: the estimate is reduced with MAX
    over the ranks BEFORE it sizes the loop.  Every search of a sharded index carries one exchange step, so a loop count taken from
    a rank's own clock desynchronises the ranks (tests/test_sharded_gloo.py::test_loop_counts_are_agreed_over_the_ranks)."""
    est = max_over_ranks(float(local_seconds_per_unit))
    return int(min(hi, max(lo, np.ceil(target_s / max(est, 1e-5)))))


def host_queries(oracle, n, d, salt=0):
    return oracle.synth_rows(SEED_QUERIES + salt, 0, n, d, True, False)


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle port; the reference is Rust and
# cannot be compiled in this image), all host threads, same config / metric / unit.
# ------------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    from oracle import oracle
    n, d, _, _, k = WORKLOADS[args.workload]
    if rank != 0:
        return
    threads = oracle.max_threads()
    # bounded sample: time each step on the first `sample_rows` rows of the same synthetic matrix and scale
    # linearly in N (the scan is linear; the sort term is O(N log N) and is scaled as such below).
    calib_rows = min(n, 100_000)
    rows = oracle.synth_rows(SEED_ROWS, 0, calib_rows, d, True, False)
    qs = host_queries(oracle, args.steps + args.warmup + 1, d)
    v = oracle.RefVecs(rows)
    t = time.perf_counter(); v.top_k_mt(qs[0], k, threads); per_row = (time.perf_counter() - t) / calib_rows
    budget_s = 90.0
    sample_rows = int(min(n, max(calib_rows, budget_s / max(args.steps + args.warmup, 1) / per_row)))
    if sample_rows > calib_rows:
        v.close()
        rows = oracle.synth_rows(SEED_ROWS, 0, sample_rows, d, True, False)
        v = oracle.RefVecs(rows)
    for i in range(args.warmup):
        v.top_k_mt(qs[i], k, threads)
    t0 = time.perf_counter()
    for i in range(args.steps):
        v.top_k_mt(qs[args.warmup + i], k, threads)
    dt = time.perf_counter() - t0
    v.close()
    scale = n / sample_rows        # the scan is linear in N; the sort's extra log factor (<= 1.1x on its ~25% share) is ignored, which favours the CPU
    step_s = dt / args.steps * scale
    qps = 1.0 / step_s
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n, d, k, args.gpus),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "cores_effective": threads,
                         "cores_affinity": oracle.affinity_threads(), "kind": "port",
                         "sample": f"{args.steps} queries over the first {sample_rows} of {n} rows, scaled x{scale:.2f} to N "
                                   f"(ref-parallel: per-row heap Vec, 3-FMA AVX2 cosine, full parallel sort, truncate)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n, d, k, world):
    shard_bytes = (n // world) * d * 4
    return {"workload": f"{args.workload.upper()}: {n} x {d} f32 resident matrix, batch-1 query, top-{k} cosine, exact-order scan",
            "n_rows": n, "dim": d, "k": k, "batch": 1, "metric": "cosine",
            "parallelism": f"row-sharded x{world} (strong scaling), per-shard top-k exchanged once per query (fused NVLink peer-memory kernel, NCCL all-gather as fallback)" if world > 1 else "single GPU",
            "l2": f"inputs larger than L2 (shard {shard_bytes / 2**20:.0f} MiB vs 126 MiB L2); a different query every step",
            "seeds": {"rows": hex(SEED_ROWS), "queries": hex(SEED_QUERIES)}}


# ------------------------------------------------------------------------------------------------------
# helpers shared by the headline and the batched configs
# ------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge
        from oracle import oracle        # cpu_baseline leg, query generation and the parity checks only
        self.torch, self.dist, self.oracle = torch, dist, oracle
        self.cg = ge.load_package()
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.uid_bytes = None
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = measured_peaks()
        self.sampler = ClockSampler(local_rank)
        self.sampler.start()

    def new_uid(self):
        """A fresh NCCL unique id per index (rank 0 creates it, everybody receives it)."""
        if self.world == 1:
            return None
        torch, dist = self.torch, self.dist
        buf = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(self.cg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def make_index(self, n, d, dtype):
        cg = self.cg
        begin, end = cg.shard_range(n, self.world, self.rank)
        dt = cg.F32 if dtype == "f32" else cg.F16
        if self.world > 1:
            ix = cg.Index(d, dt, device=self.local_rank, rank=self.rank, world=self.world, nccl_unique_id=self.new_uid(), row_offset=begin)
        else:
            ix = cg.Index(d, dt, device=self.local_rank)
        ix.reserve(end - begin)
        ix.fill_synthetic(end - begin, SEED_ROWS, True)
        for key, val in (self.args.opt or []):
            ix.set_option(key, val)
        return ix, begin, end

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def oracle_topk_over_shard(env, ix, begin, end, queries, k, chunk_rows=1_000_000):
    """CPU oracle top-k of `queries` over THIS rank's shard: rows are read back from the device in chunks (f16 widened
    exactly), every chunk is scored by the oracle (cg_fair_top_k_search_multi = cg_parallel_top_k_search's outputs) and
    the per-chunk lists are merged under the result contract.  -> list of (global rows, scores) per query."""
    oracle = env.oracle
    parts = [[] for _ in range(len(queries))]
    n_local = end - begin
    for c0 in range(0, n_local, chunk_rows):
        m = min(chunk_rows, n_local - c0)
        rows = ix.get_rows(c0, m)
        res = oracle.fair_top_k_multi(queries, rows, k)
        for qi, (ri, rs) in enumerate(res):
            parts[qi].append((ri + np.uint64(begin + c0), rs))
        del rows
    return [oracle.merge_top_k(p, k) for p in parts]


def parity_full(env, ix, begin, end, queries, k, got_rows, got_scores):
    """Full CPU-oracle check: every rank runs the oracle over its own shard, the per-shard oracle lists are gathered and
    merged (top-k of a union = top-k of the per-part top-k's), rank 0 compares with what the GPU path returned."""
    t0 = time.perf_counter()
    local = oracle_topk_over_shard(env, ix, begin, end, queries, k)
    gathered = env.gather_objects(local)
    ok, recall = True, []
    if env.rank == 0:
        for qi in range(len(queries)):
            wi, ws = env.oracle.merge_top_k([g[qi] for g in gathered], k)
            same = bool(np.array_equal(wi, got_rows[qi][:len(wi)]) and ws.tobytes() == np.ascontiguousarray(got_scores[qi][:len(ws)]).tobytes())
            ok = ok and same
            recall.append(len(set(wi.tolist()) & set(got_rows[qi].tolist())) / max(len(wi), 1))
    return {"ok": ok, "method": "oracle_full: cg_parallel_top_k_search arithmetic over ALL rows read back from every rank's device, merged on rank 0; indices and score bytes equal",
            "queries": len(queries), "recall": recall, "seconds": round(time.perf_counter() - t0, 1)}


def parity_sampled(env, ix, begin, end, queries, k, got_rows, got_scores, d_queries, sample_rows=1_000_000):
    """Shards too large for a full CPU pass inside a bench run: (1) the exact-order kernel (bit-exact against the CPU oracle
    on the full C2 matrix in this same run) must return the same lists; (2) the CPU oracle re-scores the returned rows that
    live on this rank and must reproduce their score bytes; (3) the CPU oracle scans a sample of each shard (the block that
    holds each query's best local hit plus evenly spaced blocks) and must find no row that outranks the returned k-th."""
    torch, cg, oracle = env.torch, env.cg, env.oracle
    t0 = time.perf_counter()
    nq = len(queries)
    e_rows = torch.empty((nq, k), dtype=torch.int64, device=env.dev)
    e_scores = torch.empty((nq, k), dtype=torch.float32, device=env.dev)
    e_counts = torch.empty((nq,), dtype=torch.int32, device=env.dev)
    ix.search_device(d_queries.data_ptr(), nq, k, e_rows.data_ptr(), e_scores.data_ptr(), e_counts.data_ptr(), cg.COSINE, 0, cg.PATH_EXACT)
    torch.cuda.synchronize()
    ex_r = e_rows.cpu().numpy().astype(np.uint64); ex_s = e_scores.cpu().numpy()
    same_as_exact = bool(np.array_equal(ex_r, got_rows) and ex_s.tobytes() == np.ascontiguousarray(got_scores).tobytes())
    # (2) oracle scores of the returned rows held by this rank
    rescored_ok = True
    n_local = end - begin
    for qi in range(nq):
        for j in range(k):
            g = int(got_rows[qi][j])
            if begin <= g < end:
                row = ix.get_rows(g - begin, 1)
                sc = oracle.scores(queries[qi], row)[0]
                if np.float32(sc).tobytes() != np.float32(got_scores[qi][j]).tobytes():
                    rescored_ok = False
    # (3) oracle over a sample of the shard
    blocks = 8
    per = max(min(sample_rows // blocks, n_local // blocks), 1)
    outranked = 0
    starts = sorted({min(max(0, (n_local // blocks) * b), max(n_local - per, 0)) for b in range(blocks)})
    for s0 in starts:
        m = min(per, n_local - s0)
        rows = ix.get_rows(s0, m)
        res = oracle.fair_top_k_multi(queries, rows, 1)
        for qi, (ri, rs) in enumerate(res):
            if len(ri) == 0:
                continue
            g = int(ri[0]) + begin + s0
            kth_s, kth_r = float(got_scores[qi][k - 1]), int(got_rows[qi][k - 1])
            better = (rs[0] > kth_s) or (rs[0] == kth_s and g < kth_r)
            if better and g not in set(int(x) for x in got_rows[qi]):
                outranked += 1
    recall = [len(set(ex_r[qi].tolist()) & set(int(x) for x in got_rows[qi])) / k for qi in range(nq)]
    flags = env.gather_objects({"rescored_ok": rescored_ok, "outranked": outranked})
    ok = same_as_exact and all(f["rescored_ok"] for f in flags) and sum(f["outranked"] for f in flags) == 0
    return {"ok": bool(ok), "method": "oracle_sampled+exact_kernel: shard too large for a full CPU pass in a bench run; (1) equals the exact-order "
                                      "kernel's lists (indices + score bytes), (2) CPU oracle reproduces the returned rows' score bytes, "
                                      "(3) CPU oracle over sampled row blocks of every shard finds no outranking row",
            "queries": nq, "same_as_exact_kernel": same_as_exact, "rows_sampled_per_rank": int(per * len(starts)), "recall": recall,
            "seconds": round(time.perf_counter() - t0, 1)}


# ------------------------------------------------------------------------------------------------------
# batched configs (C3 / C4 / C5)
# ------------------------------------------------------------------------------------------------------
def run_batched(env, name):
    torch, cg, oracle = env.torch, env.cg, env.oracle
    cfg = BATCHED[name]
    n, d, nq = cfg["rows"], cfg["dim"], cfg["batch"]
    esize = 4 if cfg["dtype"] == "f32" else 2
    begin, end = cg.shard_range(n, env.world, env.rank)
    need = (end - begin) * (d * esize + 4) + (3 << 30)
    free, _ = torch.cuda.mem_get_info()
    short = env.max_over_ranks(1.0 if free < need else 0.0)
    if short:
        return {"skipped": f"needs {need / 2**30:.0f} GiB per GPU, {free / 2**30:.0f} GiB free"}
    ix, begin, end = env.make_index(n, d, cfg["dtype"])
    local_rows = end - begin
    out = {"workload": f"{name.upper()}: {cfg['what']}", "n_rows": n, "dim": d, "dtype": cfg["dtype"], "batch": nq, "n_gpus": env.world,
           "scaling": "strong", "rows_per_gpu": local_rows, "runs": []}
    try:
        sample_q = sorted({0, nq // 3, (2 * nq) // 3, nq - 1})
        for k in cfg["ks"]:
            # queries: synthetic unit vectors; the sampled ones are PLANTED (a stored row + small noise), so their nearest
            # neighbour is known a priori
            nb_max = 64
            qs_host = host_queries(oracle, nq * 4, d, salt={'c3': 3, 'c4': 4, 'c5': 5}[name] * 1000 + k).reshape(4, nq, d).copy()
            planted = {}
            rng = np.random.default_rng(1234 + k)
            for qi in sample_q:
                r = int(rng.integers(0, n))
                row = oracle.synth_rows(SEED_ROWS, r, 1, d, True, cfg["dtype"] == "f16")[0]
                qs_host[0, qi] = (row + rng.normal(0.0, 0.05 / np.sqrt(d), d)).astype(np.float32)
                planted[qi] = r
            qs_dev = torch.from_numpy(qs_host).to(env.dev)
            o_r = torch.empty((nq, k), dtype=torch.int64, device=env.dev)
            o_s = torch.empty((nq, k), dtype=torch.float32, device=env.dev)
            o_c = torch.empty((nq,), dtype=torch.int32, device=env.dev)
            stream = torch.cuda.Stream(device=env.dev)

            def run(i):
                ix.search_device(qs_dev[i % 4].data_ptr(), nq, k, o_r.data_ptr(), o_s.data_ptr(), o_c.data_ptr(), cg.COSINE,
                                 stream.cuda_stream, cg.PATH_TENSOR)

            # warm-up (>= 3 batches) doubles as the estimate that sizes the timed region to ~0.4 s
            env.barrier()
            tw = time.perf_counter()
            for i in range(3):
                run(i)
            env.barrier()
            est = env.max_over_ranks((time.perf_counter() - tw) / 3)
            nb = int(min(nb_max, max(5, 0.4 / max(est, 1e-4))))
            st0 = ix.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_wall0 = time.perf_counter()
            e0.record(stream)
            for i in range(nb):
                run(i + 1)
            e1.record(stream)
            env.barrier()
            t_wall1 = time.perf_counter()
            ms = env.max_over_ranks(e0.elapsed_time(e1)) / nb
            st1 = ix.stats()
            # dominant kernel (main range of the tensor scan): CUDA events recorded by the library around that launch
            ix.set_option("reset_timing", 1); ix.set_option("timing", 1)
            for i in range(min(nb, 8)):
                run(i)
            env.barrier()
            st2 = ix.stats()
            ix.set_option("timing", 0)
            main_ms = env.max_over_ranks(st2.tc_main_ms_total / max(st2.tc_main_timed, 1))
            # results of batch 0 (the planted one) for the parity check
            run(0)
            torch.cuda.synchronize()
            got_r = o_r.cpu().numpy().astype(np.uint64); got_s = o_s.cpu().numpy()
            qsel = np.ascontiguousarray(qs_host[0][sample_q])
            gr, gs = got_r[sample_q], got_s[sample_q]
            if local_rows <= FULL_ORACLE_MAX_LOCAL_ROWS:
                parity = parity_full(env, ix, begin, end, qsel, k, gr, gs)
            else:
                parity = parity_sampled(env, ix, begin, end, qsel, k, gr, gs, torch.from_numpy(qsel).to(env.dev))
            parity["planted_nearest_neighbour_found"] = bool(all(int(gr[j][0]) == planted[qi] for j, qi in enumerate(sample_q)))
            parity["ok"] = bool(parity["ok"] and parity["planted_nearest_neighbour_found"])
            # end to end from HOST buffers
            if cfg["streaming"]:
                qstream = ix.stream(nq, k, cg.COSINE, cg.PATH_TENSOR)
                qstream.submit(qs_host[0]); qstream.submit(qs_host[1])
                qstream.flush()
                env.barrier()
                t0 = time.perf_counter()
                last = None
                for i in range(nb):
                    r = qstream.submit(qs_host[i % 4])
                    last = r if r is not None else last
                last = qstream.flush() or last
                e2e_s = env.max_over_ranks(time.perf_counter() - t0)
                qstream.submit(qs_host[0])
                last = qstream.flush()
                qstream.close()
                e2e_note = "cgvec_stream_submit / flush: double-buffered pinned upload of the next batch behind the scan of the current one, results to host"
                same = bool(last is not None and np.array_equal(last[0][sample_q], got_r[sample_q]) and last[1][sample_q].tobytes() == got_s[sample_q].tobytes())
            else:
                bufs = ix.make_search_buffers(nq, k)
                ix.search_into(qs_host[0], bufs)
                env.barrier()
                t0 = time.perf_counter()
                for i in range(nb):
                    ix.search_into(qs_host[i % 4], bufs)
                e2e_s = env.max_over_ranks(time.perf_counter() - t0)
                e2e_note = "cgvec_search_ex with host buffers: H2D of the batch + scan + exchange + results to host + sync, every batch"
                ix.search_into(qs_host[0], bufs)
                same = bool(np.array_equal(bufs["rows"][sample_q], got_r[sample_q]) and bufs["scores"][sample_q].tobytes() == got_s[sample_q].tobytes())
            env.barrier()
            alg_bytes = local_rows * d * esize                       # one pass over the local shard per tensor batch of <= 256 queries
            passes = max(1, round((st1.tc_batches - st0.tc_batches) / nb))
            flops = 2.0 * nq * local_rows * d
            pk = env.peaks
            hbm_achieved = alg_bytes / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
            tf = flops / (ms * 1e-3) / 1e12
            tensor_peak = pk["bf16_tflops_sustained"] * (0.5 if cfg["dtype"] == "f32" else 1.0)
            bound = "hbm" if name == "c4" else "tensor"
            rec = {
                "k": k, "batches_timed": nb, "ms_per_batch": ms, "value": nq / (ms * 1e-3), "unit": "queries/s",
                "roofline": {
                    "bound": bound,
                    "achieved": hbm_achieved if bound == "hbm" else tf, "peak": pk["hbm_gbs"] if bound == "hbm" else tensor_peak,
                    "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                    "frac": (hbm_achieved / pk["hbm_gbs"]) if bound == "hbm" else tf / tensor_peak,
                    "hbm": {"kernel": "tc2_scan_kernel" if nq > 128 else "tc_scan_kernel (main range)", "kernel_ms": main_ms, "achieved_gbs": hbm_achieved,
                            "frac": hbm_achieved / pk["hbm_gbs"], "frac_on_batch_time": passes * alg_bytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                            "algorithmic_bytes_per_pass_per_gpu": alg_bytes, "passes_per_batch": passes},
                    "tensor": {"tflops_per_gpu_on_batch_time": tf, "peak_tflops": tensor_peak, "frac": tf / tensor_peak,
                               "peak_note": "cuBLAS bf16 sustained (MEASURED_PEAKS.json)" + ("; halved for kind::tf32" if cfg["dtype"] == "f32" else "")},
                    "traffic": None, "peak_source": pk["source"]},
                "tc_fallbacks": int(st1.tc_fallbacks - st0.tc_fallbacks), "tc_batches": int(st1.tc_batches - st0.tc_batches),
                "gpu_launches_per_batch": (st1.kernel_launches - st0.kernel_launches) / nb,
                "exchange": {0: "none", 1: "p2p", 2: "nccl"}.get(int(st1.exchange_mode), "?"),
                "parity_ok": parity["ok"], "parity": parity,
                "e2e": {"value": nb * nq / e2e_s, "unit": "queries/s", "h2d_bytes_per_batch": nq * d * 4, "d2h_bytes_per_batch": nq * (k * 12 + 4),
                        "how": e2e_note, "same_result_as_device_path": same},
                "clocks": env.sampler.summary(t_wall0, t_wall1),
            }
            if cfg["streaming"]:
                rec["recall_at_10"] = float(np.mean(parity["recall"])) if env.rank == 0 and parity.get("recall") else None
                rec["recall_reference"] = "CPU oracle over all rows" if parity["method"].startswith("oracle_full") else "exact-order kernel (itself oracle-checked on the full C2 matrix in this run) + CPU oracle on sampled rows"
            out["runs"].append(rec)
            del qs_dev, o_r, o_s, o_c
        first = out["runs"][0]
        out.update({"value": first["value"], "unit": "queries/s", "ms_per_batch": first["ms_per_batch"], "k": first["k"],
                    "parity_ok": bool(all(r["parity_ok"] for r in out["runs"])), "tc_fallbacks": sum(r["tc_fallbacks"] for r in out["runs"]),
                    "roofline": first["roofline"]})
    finally:
        try:
            if "qstream" in locals() and qstream is not None:   # an exception between open and close: the index refuses to go while a stream is open
                qstream.close()
        except Exception:
            pass
        ix.close()
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    env = Env(args, rank, world, local_rank)
    torch, cg, oracle = env.torch, env.cg, env.oracle
    dev = env.dev

    n, d, _, _, k = WORKLOADS[args.workload]
    ix, begin, end = env.make_index(n, d, "f32")

    total = args.steps + args.warmup
    qs_host = host_queries(oracle, total, d)
    qs_dev = torch.from_numpy(qs_host).to(dev)
    out_rows = torch.empty((total, k), dtype=torch.int64, device=dev)
    out_scores = torch.empty((total, k), dtype=torch.float32, device=dev)
    out_counts = torch.empty((total,), dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def step_device(i):
        ix.search_device(qs_dev[i].data_ptr(), 1, k, out_rows[i].data_ptr(), out_scores[i].data_ptr(), out_counts[i].data_ptr(),
                         cg.COSINE, stream.cuda_stream)

    # ---- pre-warm: at least 200 ms of the same steps whatever --warmup says (clocks, L2 / TLB state, NCCL channels), so that
    #      a 20-step driver run measures the steady state and not the first two milliseconds after idle
    #      The number of rounds is agreed over the ranks (max of one timed round): every rank must issue the SAME number of
    #      searches, each of them carries one exchange step (a per-rank wall-clock loop ran different counts on different ranks
    #      and the surplus steps timed out waiting for peers that had moved on — found at 8 GPUs in r02).
    with torch.cuda.stream(stream):
        for i in range(args.warmup):                             # first round: module load, attribute calls, scratch allocation
            step_device(i)
        torch.cuda.synchronize()
        env.barrier()
        t_pre = time.perf_counter()
        for i in range(args.warmup):
            step_device(i)
        torch.cuda.synchronize()
        rounds = agreed_count(env.max_over_ranks, time.perf_counter() - t_pre, 0.25)
        for _ in range(rounds):
            for i in range(args.warmup):
                step_device(i)
            torch.cuda.synchronize()
    env.barrier()

    # ---- device-resident throughput (`value`): K back-to-back batch-1 searches on one stream, nothing else ----
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step_device(i)
    env.barrier()
    launches0 = ix.stats().kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for i in range(args.warmup, total):
            step_device(i)
        e1.record(stream)
    env.barrier()
    t_wall1 = time.perf_counter()
    dev_ms = env.max_over_ranks(e0.elapsed_time(e1))
    launches = ix.stats().kernel_launches - launches0
    clocks = env.sampler.summary(t_wall0, t_wall1)

    # ---- per-launch duration of the dominant kernel: the same steps again with a CUDA-event pair recorded by the
    #      library around every scan launch on the launch stream (kept out of the region above because an event
    #      between two kernels defeats the programmatic dependent launch that overlaps the merge with the next scan)
    ix.set_option("reset_timing", 1)
    ix.set_option("timing", 1)
    with torch.cuda.stream(stream):
        for i in range(args.warmup, total):
            step_device(i)
    env.barrier()
    st = ix.stats()
    ix.set_option("timing", 0)
    scan_ms = env.max_over_ranks(st.scan_ms_total / max(st.scans_timed, 1))
    exchange = {0: "none", 1: "p2p", 2: "nccl"}.get(int(st.exchange_mode), "?")

    # ---- what the timed region produced for its last query (checked against the oracle below, at every N) ----
    got_rows = out_rows[total - 1].cpu().numpy().astype(np.uint64)
    got_scores = out_scores[total - 1].cpu().numpy()

    # ---- end to end through the C ABI with HOST buffers: H2D query + D2H results + sync every step ----
    bufs = ix.make_search_buffers(1, k)
    q_rows = [np.ascontiguousarray(qs_host[i:i + 1]) for i in range(total)]      # one C-contiguous [1, d] view per step
    for i in range(min(args.warmup, 5)):
        ix.search_into(q_rows[i], bufs)
    env.barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, total):
        ix.search_into(q_rows[i], bufs)                      # H2D query + scan + merge/exchange + results to host + sync
    torch.cuda.synchronize()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    env.barrier()
    r_h, s_h = bufs["rows"], bufs["scores"]
    same = bool(np.array_equal(r_h[0], got_rows) and np.array_equal(s_h[0], got_scores))

    # ---- resident session (cgvec_serve_*): the same batch-1 queries served by a kernel that stays on the GPU ----
    session = None
    try:
        if args.no_extras:
            raise RuntimeError("skipped (--no-extras)")
        sess = cg.ServeSession(ix, k)
        for i in range(min(args.warmup, 5)):
            sess.search_raw(q_rows[i][0])
        env.barrier()
        t0 = time.perf_counter()
        for i in range(args.warmup, total):
            sess.search_raw(q_rows[i][0])                    # descriptor + doorbell in pinned memory, spin on the completion word
        sess_e2e_s = env.max_over_ranks(time.perf_counter() - t0)
        sess_same = bool(np.array_equal(sess.rows, got_rows) and np.array_equal(sess.scores, got_scores))
        env.barrier()
        tk = 0
        for i in range(min(args.warmup, 5)):
            tk = sess.submit_device(qs_dev[i].data_ptr(), out_rows[i].data_ptr(), out_scores[i].data_ptr(), out_counts[i].data_ptr())
        sess.wait(tk)
        ptrs = [(qs_dev[i].data_ptr(), out_rows[i].data_ptr(), out_scores[i].data_ptr(), out_counts[i].data_ptr()) for i in range(args.warmup, total)]
        env.barrier()
        sess.timer_start()                                   # CUDA event on the session's launch stream; the kernel launches inside the bracket
        for a, b_, c_, d_ in ptrs:                           # queries and results in HBM, up to 6 tickets in flight
            sess.submit_device(a, b_, c_, d_)
        sess_dev_ms = env.max_over_ranks(sess.timer_stop())  # ... drains, makes the kernel leave, second event behind it
        sess_dev_s = sess_dev_ms * 1e-3
        sess_stats = sess.stats()
        sess.close()
        torch.cuda.synchronize()
        sess_same = sess_same and bool(np.array_equal(out_rows[total - 1].cpu().numpy().astype(np.uint64), got_rows))
        session = {"device_resident": {"value": args.steps / sess_dev_s, "unit": "queries/s", "us_per_query": sess_dev_s / args.steps * 1e6},
                   "e2e": {"value": args.steps / sess_e2e_s, "unit": "queries/s", "us_per_query": sess_e2e_s / args.steps * 1e6},
                   "kernel_launches": sess_stats["launches"], "queries_served": sess_stats["served"], "kernel_launches_in_timed_region": 2,
                   "same_result_as_launch_path": sess_same,
                   "device_ms": sess_dev_ms,
                   "timing": "device_resident: CUDA events on the session's launch stream around kernel launch + K individually submitted queries + kernel exit, max over ranks; e2e: host wall clock"}
    except Exception as e:
        session = {"error": repr(e)[:300]}
        try:
            sess.close()                                     # an open session would keep the index (and every SM) busy
        except Exception:
            pass
    env.barrier()

    # ---- concurrent batch-1 callers (multi_vector_search's shape, search.rs:347-361): 16 host threads, group commit on / off ----
    concurrent = None
    if world == 1 and not args.no_extras:
        try:
            import threading
            nthreads, per = 16, max(8, min(64, args.steps // 8))
            def run_threads():
                bufs_t = [ix.make_search_buffers(1, k) for _ in range(nthreads)]
                start = threading.Barrier(nthreads + 1)
                def work(t):
                    start.wait()
                    for r in range(per):
                        ix.search_into(q_rows[(t * per + r) % total], bufs_t[t])
                th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
                [t.start() for t in th]
                start.wait(); t0 = time.perf_counter()
                [t.join() for t in th]
                dt = time.perf_counter() - t0
                last = (nthreads - 1) * per + per - 1
                return nthreads * per / dt, bufs_t[nthreads - 1]["rows"][0].copy(), last % total
            ix.set_option("coalesce", 1)
            run_threads()
            st0 = ix.stats()
            qps_on, rows_on, qi = run_threads()
            st1 = ix.stats()
            ix.set_option("coalesce", 0)
            qps_off, rows_off, _ = run_threads()
            ix.set_option("coalesce", 1)
            ix.search_into(q_rows[qi], bufs)
            concurrent = {"threads": nthreads, "queries_per_thread": per, "value": qps_on, "unit": "queries/s",
                          "without_group_commit": qps_off, "single_caller": args.steps / e2e_s,
                          "coalesced_batches": int(st1.coalesced_batches - st0.coalesced_batches),
                          "coalesced_queries": int(st1.coalesced_queries - st0.coalesced_queries),
                          "same_result_as_single_caller": bool(np.array_equal(rows_on, bufs["rows"][0]) and np.array_equal(rows_off, bufs["rows"][0]))}
        except Exception as e:
            concurrent = {"error": repr(e)[:300]}

    # ---- parity at every N: the CPU oracle over every rank's shard (rows read back from the device), merged on rank 0 ----
    par = parity_full(env, ix, begin, end, np.ascontiguousarray(qs_host[total - 1:total]), k, got_rows[None, :], got_scores[None, :])

    # `value`: the faster of the library's two batch-1 transports on this shard size, both device-timed with CUDA events over the
    # same K individually submitted queries (both are in the line).  The launch-per-query chain wins on long scans, the resident
    # session on short ones (per-GPU shards at 4-8 GPUs), where the ~9 us per launch it removes matter.
    launch_dev_ms = dev_ms
    transport = "launch per query (programmatic launch chain)"
    if session and "device_ms" in session and session.get("same_result_as_launch_path") and session["device_ms"] < dev_ms:
        dev_ms = session["device_ms"]
        launches = session["kernel_launches_in_timed_region"]
        transport = "resident session (cgvec_serve_*)"
    qps = args.steps / (dev_ms * 1e-3)
    e2e_qps = args.steps / e2e_s
    local_rows = end - begin
    alg_bytes = local_rows * d * 4 + local_rows * 4            # matrix once + one f32 norm per row
    peak, peak_src = env.peaks["hbm_gbs"], env.peaks["source"]
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            tj = json.load(f)
        if world == 1:         # the ncu capture is of the 1-GPU launch; at N > 1 the shard (and the traffic) is 1/N of it
            traffic = tj.get("dram_bytes_per_launch")
            traffic_note = "dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture (profiles/), not measured in this run"
        else:              # scaled by the shard's share of the rows (the kernel reads every local row exactly once)
            traffic = tj.get("dram_bytes_per_launch") * local_rows / n if tj.get("dram_bytes_per_launch") else None
            traffic_note = "the committed 1-GPU ncu capture scaled by this rank's share of the rows; not measured in this run"
    except Exception:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = oracle.max_threads()
        rows_host = oracle.synth_rows(SEED_ROWS, 0, n, d, True, False)
        # labelled NON-reference (BASELINE.md §2 "fair-cpu"): contiguous matrix, per-thread bounded selection, no full sort
        oracle.fair_top_k_mt(qs_host[0], rows_host, k, threads)
        nf, t0 = 0, time.perf_counter()
        while nf < 5 or (time.perf_counter() - t0 < 3.0 and nf < 100):
            oracle.fair_top_k_mt(qs_host[nf % total], rows_host, k, threads)
            nf += 1
        fair_qps = nf / (time.perf_counter() - t0)
        v = oracle.RefVecs(rows_host)
        del rows_host
        v.top_k_mt(qs_host[0], k, threads)
        nq_cpu, t0 = 0, time.perf_counter()
        while nq_cpu < 8 or (time.perf_counter() - t0 < 10.0 and nq_cpu < 200):
            v.top_k_mt(qs_host[nq_cpu % total], k, threads)
            nq_cpu += 1
        cpu_s = (time.perf_counter() - t0) / nq_cpu
        v.close()
        cpu_baseline = {"value": 1.0 / cpu_s, "unit": "queries/s", "cores": threads, "cores_effective": threads,
                        "cores_affinity": oracle.affinity_threads(), "kind": "port",
                        "sample": f"{nq_cpu} batch-1 queries over the full {n} x {d} matrix (ref-parallel port of simd_ops.rs:361-383: "
                                  f"per-row heap Vec, 3-FMA AVX2 cosine, full parallel sort, truncate); threads = min(affinity, cgroup cpu quota)",
                        "gpu_matches_oracle_on_full_matrix": par["ok"],
                        "fair_cpu_non_reference": {"value": fair_qps, "unit": "queries/s",
                                                   "what": "same arithmetic, contiguous matrix, per-thread bounded top-k instead of the reference's per-row heap Vec + full parallel sort"}}
    ix.close()
    torch.cuda.empty_cache()

    configs = {}
    wanted = [c for c in (args.configs.split(",") if args.configs and args.configs != "none" else []) if c in BATCHED]
    for name in wanted:
        try:
            configs[name] = run_batched(env, name)
        except Exception as e:      # a failing extra config must not take the headline line with it
            configs[name] = {"error": repr(e)[:300]}
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass
    env.sampler.stop_flag.set()

    if rank == 0:
        step_ms = dev_ms / args.steps
        roof_kernel, roof_kernel_ms, roof_bytes, roof_note = "scan_exact_kernel<float,COSINE,1>", scan_ms, alg_bytes, None
        if transport.startswith("resident"):
            # one launch of the resident kernel covers the whole timed region: K passes over the shard
            roof_kernel, roof_kernel_ms, roof_bytes = "scan_serve_kernel<float,COSINE>", dev_ms, alg_bytes * args.steps
            achieved = roof_bytes / (dev_ms * 1e-3) / 1e9
            if traffic:
                traffic, traffic_note = traffic * args.steps, (traffic_note or "") + "; times the K passes of the one resident launch"
            roof_note = {"launch_path_kernel": "scan_exact_kernel<float,COSINE,1>", "launch_path_kernel_ms": scan_ms,
                         "launch_path_frac": alg_bytes / (scan_ms * 1e-3) / 1e9 / peak if scan_ms > 0 else None}
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n, d, k, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_on_step_time": alg_bytes / (step_ms * 1e-3) / 1e9 / peak if step_ms > 0 else None,
                         "traffic": traffic, "traffic_note": traffic_note, "kernel": roof_kernel, "kernel_ms": roof_kernel_ms,
                         "algorithmic_bytes_per_launch": roof_bytes, "peak_source": peak_src, "other_transport": roof_note,
                         "step_share": roof_kernel_ms / dev_ms if transport.startswith("resident") else (scan_ms / step_ms if dev_ms > 0 else None),
                         "step_share_note": "above 1 when consecutive scans overlap (programmatic launch chain): the kernel timed alone is longer than a step",
                         "kernel_timing": f"{int(st.scans_timed)} launches, CUDA event pair around each on the launch stream, pass run right after the timed region",
                         "geometry": {"grid": st.grid, "block": st.block, "smem": st.smem_bytes, "stages": st.stages,
                                      "tile_rows": st.tile_rows}},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": d * 4, "d2h_bytes_per_step": k * 12 + 4,
                    "same_result_as_device_path": same},
            "parity_ok": par["ok"], "parity": par,
            "batch1_transport": transport,
            "transports": {"launch_per_query": {"value": args.steps / (launch_dev_ms * 1e-3), "ms_per_step": launch_dev_ms / args.steps},
                           "resident_session": ({"value": session["device_resident"]["value"], "ms_per_step": session["device_ms"] / args.steps}
                                                if session and "device_ms" in session else None)},
            "exchange": exchange,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "session": session,
            "concurrent_callers": concurrent,
            "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--configs", default="c3,c4,c5", help="comma-separated batched configs to add to the line (c3,c4,c5) or 'none'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the resident-session and concurrent-caller records (profiler runs: a resident kernel must not be serialised by ncu)")
    ap.add_argument("--opt", action="append", type=lambda s: (s.split("=")[0], int(s.split("=")[1])),
                    help="library tuning knob key=value (repeatable)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        print(json.dumps({"error": f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}"}), flush=True)
        sys.exit(2)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
