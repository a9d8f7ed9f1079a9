"""Row-sharded index across 2+ GPUs (one process per GPU, NCCL all-gather of the per-shard top-k, device merge):
every rank must return exactly the single-index answer of the oracle.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, uid_q, rows, qs, k, metric_name, out):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    cg = ge.load_package()
    import torch
    torch.cuda.set_device(rank)
    if rank == 0:
        uid = cg.nccl_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    b, e = cg.shard_range(len(rows), world, rank)
    ix = cg.Index(rows.shape[1], cg.F32, device=rank, rank=rank, world=world, nccl_unique_id=uid, row_offset=b)
    ix.add(rows[b:e])
    metric = {"cosine": cg.COSINE, "l2": cg.L2}[metric_name]
    r, s, c = ix.search(qs, k, metric)
    out[rank] = (r.tolist(), s.tobytes(), c.tolist())
    ix.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("metric_name", ["cosine", "l2"])
def test_sharded_search_equals_oracle(oracle, metric_name):
    import torch.multiprocessing as mp
    world = min(_ngpus(), 4)
    rng = np.random.default_rng(31)
    n, d, k = 50_000, 256, 20
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[40_000] = rows[5]                       # tie across shards: lower global row must win
    qs = rng.standard_normal((3, d)).astype(np.float32)
    qs[0] = rows[5]
    om = oracle.COSINE if metric_name == "cosine" else oracle.L2
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict(); uid_q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, uid_q, rows, qs, k, metric_name, out)) for r in range(world)]
        [p.start() for p in procs]
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        for qi in range(3):
            wi, ws = oracle.parallel_top_k_search(qs[qi], rows, k, metric=om)
            for r in range(world):
                gr, gs, gc = out[r]
                assert gc[qi] == k
                assert gr[qi] == wi.tolist()
                assert np.frombuffer(gs, np.float32).reshape(3, k)[qi].tobytes() == ws.tobytes()
