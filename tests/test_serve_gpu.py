"""Resident batch-1 sessions (cgvec_serve_*, csrc/scan_serve.cuh): the scan kernel stays on the GPU between queries.  Every result
must equal the oracle's parallel_top_k_search (simd_ops.rs:361-383) bit for bit, like the launch-per-query path."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(oracle, got, q, ref, k, metric=None):
    rows, scores, count = got[:3]
    wi, ws = oracle.parallel_top_k_search(q, ref, k) if metric is None else oracle.parallel_top_k_search(q, ref, k, metric=metric)
    assert count == len(wi)
    assert rows[:count].tolist() == wi.tolist()
    assert scores[:count].tobytes() == ws.tobytes()


@pytest.mark.parametrize("n,d,k", [(20_000, 128, 10), (4_099, 100, 7), (300, 64, 10), (50_000, 768, 10), (9_000, 33, 32)])
def test_session_matches_oracle(cg, oracle, n, d, k):
    """Sequential queries through one resident kernel: full tiles, ragged row tails, d % 8 != 0 tails, fewer tiles than SMs."""
    rng = np.random.default_rng(n + d)
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[n // 2] = rows[3]                                       # a tie across CTAs: lower row first
    ix = cg.Index(d)
    ix.add(rows)
    s = cg.ServeSession(ix, k)
    qs = rng.standard_normal((12, d)).astype(np.float32)
    qs[0] = rows[3]
    for q in qs:
        _check(oracle, s.search(q), q, rows, k)
    st = s.stats()
    assert st["served"] == len(qs) and st["launches"] >= 1
    s.close()
    # the launch-per-query path still works afterwards and agrees
    r, sc, c = ix.search(qs[1], k)
    wi, ws = oracle.parallel_top_k_search(qs[1], rows, k)
    assert r[0].tolist() == wi.tolist() and sc[0].tobytes() == ws.tobytes()
    ix.close()


@pytest.mark.parametrize("metric_name", ["l2", "dot"])
def test_session_metrics_and_f16(cg, oracle, metric_name):
    rng = np.random.default_rng(5)
    n, d, k = 30_000, 256, 10
    rows = (rng.standard_normal((n, d)) / 8).astype(np.float32)
    ref = rows.astype(np.float16).astype(np.float32)
    ix = cg.Index(d, cg.F16)
    ix.add(rows)
    metric, om = (cg.L2, oracle.L2) if metric_name == "l2" else (cg.DOT, oracle.DOT)
    s = cg.ServeSession(ix, k, metric)
    for q in rng.standard_normal((6, d)).astype(np.float32):
        _check(oracle, s.search(q), q, ref, k, metric=om)
    s.close(); ix.close()


def test_session_ids_zero_norm_rows_and_small_k(cg, oracle):
    import uuid
    rng = np.random.default_rng(6)
    n, d = 5_000, 96
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[17] = 0.0                                               # zero norm -> cosine 0 (simd_ops.rs:73-74)
    ids = [uuid.UUID(int=i + 1) for i in range(n)]
    ix = cg.Index(d)
    ix.add(rows, ids)
    s = cg.ServeSession(ix, 1)
    q = rows[123]
    r, sc, c, gids = s.search(q, want_ids=True)
    assert c == 1 and int(r[0]) == 123 and gids[0].tobytes() == ids[123].bytes
    with pytest.raises(cg.CgvecError):
        ix.close()                                               # an open session keeps the index alive
    with pytest.raises(cg.CgvecError):
        ix.add(rows[:1])                                         # and the write side out
    s.close()
    s = cg.ServeSession(ix, 64)
    for q in rng.standard_normal((3, d)).astype(np.float32):
        _check(oracle, s.search(q), q, rows, 64)
    s.close(); ix.close()


def test_session_leaves_when_idle_and_comes_back(cg, oracle):
    """The resident grid owns every SM, so it leaves after `idle_us` without a doorbell; the next submit restarts it
    (exit announcement / doorbell race included: many short gaps around the idle time)."""
    rng = np.random.default_rng(7)
    n, d, k = 40_000, 128, 10
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ix = cg.Index(d)
    ix.add(rows)
    s = cg.ServeSession(ix, k, idle_us=100)
    qs = rng.standard_normal((40, d)).astype(np.float32)
    want = [oracle.parallel_top_k_search(q, rows, k) for q in qs]
    for i, q in enumerate(qs):
        r, sc, c = s.search(q)
        assert r.tolist() == want[i][0].tolist() and sc.tobytes() == want[i][1].tobytes()
        time.sleep([0.0, 0.00005, 0.0001, 0.0002, 0.002][i % 5])
    assert s.stats()["launches"] >= 3
    # while the session is idle other work runs on the device
    r, sc, c = ix.search(qs[0], k)
    assert r[0].tolist() == want[0][0].tolist()
    s.pause()
    r, sc, c = s.search(qs[1])
    assert r.tolist() == want[1][0].tolist()
    s.close(); ix.close()


def test_session_pipelined_device_submissions(cg, oracle):
    """Queries and results in device memory, several tickets in flight (the ring keeps streaming rows across queries)."""
    import torch
    rng = np.random.default_rng(8)
    n, d, k, nq = 60_000, 128, 10, 40
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ix = cg.Index(d)
    ix.add(rows)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    dq = torch.from_numpy(qs).cuda()
    d_rows = torch.full((nq, k), -1, dtype=torch.int64, device="cuda")
    d_scores = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros((nq,), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    s = cg.ServeSession(ix, k)
    tickets = []
    for i in range(nq):
        tickets.append(s.submit_device(dq[i].data_ptr(), d_rows[i].data_ptr(), d_scores[i].data_ptr(), d_counts[i].data_ptr()))
    s.wait(tickets[-1])
    s.pause()
    gr, gs, gc = d_rows.cpu().numpy(), d_scores.cpu().numpy(), d_counts.cpu().numpy()
    for i in range(nq):
        wi, ws = oracle.parallel_top_k_search(qs[i], rows, k)
        assert gc[i] == k and gr[i].tolist() == wi.tolist() and gs[i].tobytes() == ws.tobytes(), i
    s.close(); ix.close()


def test_session_config2_full_size(cg, oracle):
    """BASELINE config 2 (1M x 768 f32, batch-1, top-10) through a session: equal to the launch-per-query path and the oracle."""
    from tests import synth
    n, d = 1_000_000, 768
    ix = cg.Index(d)
    ix.fill_synthetic(n, 0xC0DE6A9F, True)
    qs = synth.synth_rows(0x5EED0001, list(range(8)), d)
    want = [ix.search(q, 10) for q in qs]
    s = cg.ServeSession(ix, 10)
    t0 = time.perf_counter()
    for i, q in enumerate(qs):
        r, sc, c = s.search(q)
        assert c == 10 and r.tolist() == want[i][0][0].tolist() and sc.tobytes() == want[i][1][0].tobytes()
    dt = (time.perf_counter() - t0) / len(qs)
    s.close()
    rows = ix.get_rows(0, n)
    wi, ws = oracle.parallel_top_k_search(qs[0], rows, 10)
    assert want[0][0][0].tolist() == wi.tolist() and want[0][1][0].tobytes() == ws.tobytes()
    print(f"session: {dt * 1e3:.3f} ms per query end to end")
    ix.close()
