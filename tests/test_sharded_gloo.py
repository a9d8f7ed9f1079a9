"""world_size-2 CPU test of the N>1 host path: each rank takes its cgvec_shard_range, produces its partial
top-k (the oracle stands in for the GPU scan here — this test covers partitioning, the exchange layout and
cgvec_merge_topk_host, not scoring), all-gathers the packed partials over gloo exactly as the NCCL path does
on the device, and merges.  Every rank must end with the global top-k."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, d, k, metric, q, rows, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    from oracle import oracle
    cg = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = cg.shard_range(n, world, rank)
        idx, sc = oracle.parallel_top_k_search(q, rows[b:e], k, metric=metric)
        pack = torch.zeros(2 * k + 1, dtype=torch.float64)
        pack[0] = len(idx)
        pack[1:1 + len(idx)] = torch.from_numpy((idx + b).astype(np.float64))
        pack[1 + k:1 + k + len(sc)] = torch.from_numpy(sc.astype(np.float64))
        gathered = [torch.zeros_like(pack) for _ in range(world)]
        dist.all_gather(gathered, pack)                       # the single collective of the path
        prow = np.stack([g[1:1 + k].numpy().astype(np.uint64) for g in gathered])
        psc = np.stack([g[1 + k:1 + 2 * k].numpy().astype(np.float32) for g in gathered])
        pcnt = np.array([int(g[0]) for g in gathered], np.uint32)
        mi, ms = cg.merge_topk_host(prow, psc, pcnt, k, ascending=(metric == oracle.L2))
        out[rank] = (mi.tolist(), ms.tobytes())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric_name", ["cosine", "l2"])
def test_two_rank_gloo_merge_equals_single_process(oracle, metric_name):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(21)
    n, d, k, world = 777, 40, 15, 2
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[700] = rows[3]                                      # tie across the shard boundary
    q = rng.standard_normal(d).astype(np.float32)
    metric = oracle.COSINE if metric_name == "cosine" else oracle.L2
    want_idx, want_sc = oracle.parallel_top_k_search(q, rows, k, metric=metric)
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + (os.getpid() % 1000) + (0 if metric_name == "cosine" else 1)
        procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, metric, q, rows, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        for r in range(world):
            gi, gs = out[r]
            assert gi == want_idx.tolist()
            assert gs == want_sc.tobytes()


def _worker_counts(rank, world, port, out):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def max_over_ranks(x):
            t = torch.tensor([x], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        # every rank measured a different time per round (its own clock): the loop count must still be the same everywhere
        local = [0.0013, 0.0021][rank]
        out[rank] = [bench.agreed_count(max_over_ranks, local * f, 0.25) for f in (1.0, 0.37, 40.0, 1e-9)]
    finally:
        dist.destroy_process_group()


def test_loop_counts_are_agreed_over_the_ranks():
    """bench.py sizes its pre-warm (and the batched configs' timed regions) from a measured time per unit.  On a sharded index every
    search carries an exchange step, so all ranks must issue the same number of searches: a per-rank wall-clock loop ran different
    counts at 8 GPUs in r02 and the surplus steps timed out in the exchange (profiles/r02_bench_n8_prewarm_failure.txt)."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + (os.getpid() % 1000) + 7
        procs = [ctx.Process(target=_worker_counts, args=(r, world, port, out)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert out[0] == out[1]
        assert out[0][0] == int(np.ceil(0.25 / 0.0021)) and out[0][2] == 3 and out[0][3] == 500      # sized by the SLOWEST rank, clamped
