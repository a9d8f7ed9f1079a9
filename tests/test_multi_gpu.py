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


def _worker(rank, world, uid_q, rows, qs, k, metric_name, out, dtype_name="f32", path_name="auto"):
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
    dt = cg.F32 if dtype_name == "f32" else cg.F16
    ix = cg.Index(rows.shape[1], dt, device=rank, rank=rank, world=world, nccl_unique_id=uid, row_offset=b)
    ix.add(rows[b:e])
    metric = {"cosine": cg.COSINE, "l2": cg.L2}[metric_name]
    path = {"auto": cg.PATH_AUTO, "tensor": cg.PATH_TENSOR}[path_name]
    r, s, c = ix.search(qs, k, metric, path=path)
    st = ix.stats()
    out[rank] = (r.tolist(), s.tobytes(), c.tolist(), int(st.tc_batches), int(st.tc_fallbacks))
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
                gr, gs, gc = out[r][:3]
                assert gc[qi] == k
                assert gr[qi] == wi.tolist()
                assert np.frombuffer(gs, np.float32).reshape(3, k)[qi].tobytes() == ws.tobytes()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("dtype_name", ["f32", "f16"])
def test_single_process_multi_device_index(cg, oracle, dtype_name):
    """cgvec_create(n_devices > 1): one host process drives every GPU (the Rust-server deployment).  Rows are dealt in
    1024-row blocks round-robin; results, ids, get_embedding, rescore and upserts must match a single index."""
    import uuid
    G = min(_ngpus(), 4)
    dt = cg.F32 if dtype_name == "f32" else cg.F16
    rng = np.random.default_rng(41)
    n, d = 7_777, 192
    rows = (rng.standard_normal((n, d)) / 14).astype(np.float32)
    rows[7_000] = rows[9]                                     # tie across shards
    ref = rows if dt == cg.F32 else rows.astype(np.float16).astype(np.float32)
    ids = [uuid.UUID(int=i + 1) for i in range(n)]
    ix = cg.Index(d, dt, devices=list(range(G)))
    ix.add(rows[:3000], ids[:3000]); ix.add(rows[3000:], ids[3000:])          # incremental adds cross block boundaries
    assert len(ix) == n
    qs = rng.standard_normal((6, d)).astype(np.float32)
    qs[0] = rows[9]
    for metric, om in ((cg.COSINE, oracle.COSINE), (cg.L2, oracle.L2), (cg.DOT, oracle.DOT)):
        r, s, c, got_ids = ix.search(qs, 25, metric, want_ids=True)
        for qi in range(len(qs)):
            wi, ws = oracle.parallel_top_k_search(qs[qi], ref, 25, metric=om)
            assert r[qi].tolist() == wi.tolist()
            assert s[qi].tobytes() == ws.tobytes()
            assert [uuid.UUID(bytes=got_ids[qi, j].tobytes()).int - 1 for j in range(25)] == wi.tolist()
    assert ix.get_rows(1000, 2500).tobytes() == ref[1000:3500].tobytes()
    assert ix.get(ids[5000]).tobytes() == ref[5000].tobytes()
    sel = np.array([0, 1023, 1024, 2048, 7776], np.uint64)
    assert ix.rescore(qs[1], sel, cg.COSINE, cg.FORMULA_SEQ).tobytes() == np.float32([oracle.cosine_similarity_seq(qs[1], ref[int(i)]) for i in sel]).tobytes()
    assert ix.distances_first(qs[2], 1500).tobytes() == oracle.compute_distances_cpu(qs[2], ref.reshape(-1), d, 1500).tobytes()
    ix.add(-rows[9:10], [ids[9]])                              # upsert by id
    assert len(ix) == n and ix.search(qs[0], 1)[0][0, 0] == 7_000
    # k beyond the fused peer exchange (128): per-device lists pulled into the first device and merged there.  SemanticSearch
    # over-fetches max(3*limit, limit+10) (search.rs:113), so limit 100 needs k = 300
    for kk in (129, 300, 1024):
        r, s, c = ix.search(qs, kk)
        for qi in range(len(qs)):
            wi, ws = oracle.parallel_top_k_search(qs[qi], np.vstack([ref[:9], -ref[9:10], ref[10:]]), kk)
            assert r[qi].tolist() == wi.tolist(), (kk, qi)
            assert s[qi].tobytes() == ws.tobytes(), (kk, qi)
    with pytest.raises(cg.CgvecError):
        ix.search(qs, 1025)                                    # beyond the fused top-k limit
    ix.close()
    # device-generated synthetic rows are the same matrix whatever the sharding
    a = cg.Index(d, dt, devices=list(range(G))); a.fill_synthetic(5000, 77, True)
    b = cg.Index(d, dt); b.fill_synthetic(5000, 77, True)
    assert a.get_rows(0, 5000).tobytes() == b.get_rows(0, 5000).tobytes()
    a.close(); b.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("dtype_name,k", [("f16", 100), ("f16", 300), ("f32", 10)])
def test_single_process_multi_device_tensor_and_device_io(cg, oracle, dtype_name, k):
    """The single-process multi-device index on the tensor-core path (batch-64, block-dealt shards: candidate rows are mapped
    back to shard-local rows for the exact rescore) and with device-resident queries/results on the first device."""
    import torch
    G = min(_ngpus(), 4)
    dt = cg.F32 if dtype_name == "f32" else cg.F16
    rng = np.random.default_rng(59)
    n, d, nq = 40_000 * G + 333, 256, 64
    rows = (rng.standard_normal((n, d)) / 16).astype(np.float32)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    ref = rows if dt == cg.F32 else rows.astype(np.float16).astype(np.float32)
    ix = cg.Index(d, dt, devices=list(range(G)))
    ix.add(rows)
    want = {qi: oracle.parallel_top_k_search(qs[qi], ref, k) for qi in range(0, nq, 9)}
    if k <= 512:
        r, s, c = ix.search(qs, k, cg.COSINE, path=cg.PATH_TENSOR)
        for qi, (wi, ws) in want.items():
            assert r[qi].tolist() == wi.tolist(), qi
            assert s[qi].tobytes() == ws.tobytes(), qi
        r2, s2, _ = ix.search(qs, k, cg.COSINE)                    # AUTO picks the tensor path on shards this size
        assert r2.tobytes() == r.tobytes() and s2.tobytes() == s.tobytes()
    # device I/O: queries and results live on the first device, the call is asynchronous on the caller's stream
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dq = torch.from_numpy(qs).cuda(non_blocking=False)
        d_rows = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        d_scores = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        d_counts = torch.empty((nq,), dtype=torch.int32, device="cuda")
        for path in (cg.PATH_EXACT, cg.PATH_AUTO):
            d_rows.fill_(-1)
            ix.search_device(dq.data_ptr(), nq, k, d_rows.data_ptr(), d_scores.data_ptr(), d_counts.data_ptr(), cg.COSINE,
                             stream=st.cuda_stream, path=path)
            st.synchronize()
            gr, gs = d_rows.cpu().numpy(), d_scores.cpu().numpy()
            assert d_counts.cpu().tolist() == [k] * nq
            for qi, (wi, ws) in want.items():
                assert gr[qi].tolist() == wi.tolist(), (path, qi)
                assert gs[qi].tobytes() == ws.tobytes(), (path, qi)
    ix.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_non_simd_formulas_on_a_multi_device_index(cg, oracle):
    """search_baseline (optimization.rs:376-402) and InMemoryVectorStore::search_similar (graph_vector.rs:479-494) on the
    reference's own N=1000, d=128, seed 11223 vectors, served by a single-process multi-device index."""
    vecs = oracle.generate_optimization_vectors(5000, 128, 11223)
    ix = cg.Index(128, devices=list(range(min(_ngpus(), 4))))
    ix.add(vecs)
    for qi in (0, 1234, 4999):
        want_i, want_d = oracle.search_baseline(vecs[qi], vecs, 10)
        r, s, c = ix.search(vecs[qi], 10, cg.COSINE, formula=cg.FORMULA_BASELINE)
        assert int(c[0]) == 10 and r[0].tolist() == want_i.tolist() and s[0].tobytes() == want_d.tobytes()
        wi, ws = oracle.inmemory_search_similar(vecs[qi], vecs, 150)
        r2, s2, _ = ix.search(vecs[qi], 150, cg.COSINE, formula=cg.FORMULA_SEQ)
        assert r2[0].tolist() == wi.tolist() and s2[0].tobytes() == ws.tobytes()
    ix.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_cpp_semantic_search_on_a_two_device_store(cg, oracle):
    """The compiled C++ SemanticSearch mirror over a 2-device B200VectorStore with limit 100 (over-fetch 300 > the fused
    exchange's k limit): the Rust-server deployment INTEGRATION.md recommends."""
    import subprocess
    demo = cg._build.build_host_demo()
    n, dim, limit = 3000, 384, 100
    out = subprocess.run([demo, str(n), str(dim), str(limit), "2"], capture_output=True, text=True, check=True).stdout
    lines = dict(l.split(" ", 1) for l in out.strip().splitlines())
    assert "error" not in lines, lines.get("error")
    embs = np.stack([oracle.hash_text_embedding(f"fn item_{i}() {{}}", dim) for i in range(n)])
    q = oracle.hash_text_embedding("fn item_42() { }", dim)
    wi, ws = oracle.parallel_top_k_search(q, embs, limit)
    assert lines["search_similar"].split() == [f"00000000-0000-0000-0000-{int(i) + 1:012x}" for i in wi]
    si, snorm, _ = oracle.search_by_embedding(q, embs, limit)
    sem = [t.rsplit(":", 1) for t in lines["semantic"].split()]
    assert [u for u, _ in sem] == [f"00000000-0000-0000-0000-{int(i) + 1:012x}" for i in si]
    assert np.float32([float.fromhex(x) for _, x in sem]).tobytes() == snorm.tobytes()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("dtype_name,k", [("f16", 100), ("f16", 10), ("f32", 100)])
def test_sharded_tensor_path_equals_oracle(oracle, dtype_name, k):
    """BASELINE config 4's shape in small: row-sharded f16 (and f32/TF32) index, batch-64 queries, tensor-core scan per
    shard, exact per-shard top-k exchanged with ONE NCCL all-gather, every rank returns the oracle's answer."""
    import torch.multiprocessing as mp
    world = min(_ngpus(), 4)
    rng = np.random.default_rng(53)
    n, d, nq = 120_000, 256, 64
    rows = (rng.standard_normal((n, d)) / 16).astype(np.float32)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    ref = rows if dtype_name == "f32" else rows.astype(np.float16).astype(np.float32)
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict(); uid_q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, uid_q, rows, qs, k, "cosine", out, dtype_name, "tensor")) for r in range(world)]
        [p.start() for p in procs]
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
        for r in range(world):
            gr, gs, gc, batches, fallbacks = out[r]
            assert batches >= 1 and fallbacks == 0
            got_s = np.frombuffer(gs, np.float32).reshape(nq, k)
            for qi in range(0, nq, 7):
                wi, ws = oracle.parallel_top_k_search(qs[qi], ref, k)
                assert gr[qi] == wi.tolist()
                assert got_s[qi].tobytes() == ws.tobytes()


def _worker_general(rank, world, uid_q, rows, qs, k, out, bounds, dtype_name, path_name, reps, skew_rank):
    """Like _worker, with explicit shard boundaries, repeated calls and a host-side delay on one rank between calls (rank
    skew: the other ranks' exchange kernels spin on this rank's flags while their next scans are already queued)."""
    sys.path.insert(0, ROOT)
    import time
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
    b, e = bounds[rank], bounds[rank + 1]
    dt = cg.F32 if dtype_name == "f32" else cg.F16
    ix = cg.Index(rows.shape[1], dt, device=rank, rank=rank, world=world, nccl_unique_id=uid, row_offset=b)
    ix.add(rows[b:e])
    path = {"auto": cg.PATH_AUTO, "tensor": cg.PATH_TENSOR, "exact": cg.PATH_EXACT}[path_name]
    res = []
    rng = np.random.default_rng(rank)
    for it in range(reps):
        if rank == skew_rank:
            time.sleep(float(rng.uniform(0.0, 0.02)))
        r, s, c = ix.search(qs[it % len(qs)], k, cg.COSINE, path=path)
        res.append((r.tolist(), s.tobytes(), c.tolist()))
    st = ix.stats()
    out[rank] = (res, int(st.tc_batches), int(st.tc_fallbacks), int(st.exchange_mode))
    ix.close()


def _run_general(oracle, rows, qsets, k, bounds, dtype_name="f32", path_name="auto", reps=1, skew_rank=-1):
    import torch.multiprocessing as mp
    world = len(bounds) - 1
    ref = rows if dtype_name == "f32" else rows.astype(np.float16).astype(np.float32)
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict(); uid_q = ctx.Queue()
        procs = [ctx.Process(target=_worker_general, args=(r, world, uid_q, rows, qsets, k, out, bounds, dtype_name, path_name, reps, skew_rank))
                 for r in range(world)]
        [p.start() for p in procs]
        for p in procs:
            p.join(600)
            assert p.exitcode == 0, "a rank failed or hung"
        want = {}
        for r in range(world):
            res = out[r][0]
            for it in range(reps):
                qs = qsets[it % len(qsets)]
                gr, gs, gc = res[it]
                got_s = np.frombuffer(gs, np.float32).reshape(len(qs), k)
                for qi in range(len(qs)):
                    key = (it % len(qsets), qi)
                    if key not in want:
                        want[key] = oracle.parallel_top_k_search(qs[qi], ref, k)
                    wi, ws = want[key]
                    assert gr[qi][:len(wi)] == wi.tolist(), (r, it, qi)
                    assert got_s[qi][:len(ws)].tobytes() == ws.tobytes(), (r, it, qi)
        return {r: out[r][1:] for r in range(world)}


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_tiny_shards_many_queries_under_rank_skew(oracle):
    """Shards far smaller than one tile per SM (the scan grid does not fill the GPU, so it never releases its dependents
    early), 14 queries per call (four chained scan -> exchange steps) and one rank that dawdles between calls: the
    double-buffered per-CTA lists and the two-parity exchange slots must never be overwritten early."""
    world = min(_ngpus(), 4)
    rng = np.random.default_rng(77)
    n, d, k = 1_500 * world, 128, 10
    rows = rng.standard_normal((n, d)).astype(np.float32)
    qsets = [rng.standard_normal((14, d)).astype(np.float32) for _ in range(3)]
    bounds = [n * r // world for r in range(world + 1)]
    st = _run_general(oracle, rows, qsets, k, bounds, "f32", "exact", reps=40, skew_rank=world - 1)
    assert all(v[2] == 1 for v in st.values()), "the fused peer-memory exchange should have carried these calls"


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_uneven_shards_take_the_same_path_on_every_rank(oracle):
    """Rank 0 holds enough rows for the tensor path under AUTO, the last rank does not: the decision must be taken on a
    rank-invariant quantity or the ranks issue different collectives and hang."""
    world = 2
    rng = np.random.default_rng(78)
    n, d, k = 70_000, 128, 10
    rows = (rng.standard_normal((n, d)) / 11).astype(np.float32)
    qsets = [rng.standard_normal((64, d)).astype(np.float32)]
    st = _run_general(oracle, rows, qsets, k, [0, 40_000, 70_000], "f16", "auto", reps=2)
    assert len({v[0] > 0 for v in st.values()}) == 1, "ranks disagreed on the kernel family"


def _worker_session(rank, world, uid_q, rows, qs, k, out, skew_rank):
    """Every rank opens a resident session on its shard and submits the same queries in the same order; the CTA that finishes a
    query on each rank runs the NVLink peer exchange inside the resident kernel."""
    sys.path.insert(0, ROOT)
    import time
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
    r0, s0, _ = ix.search(qs[0], k)                              # launch-per-query path first (shares the exchange buffers and counter)
    sess = cg.ServeSession(ix, k, idle_us=300)
    res = []
    rng = np.random.default_rng(rank)
    for i, q in enumerate(qs):
        if rank == skew_rank and i % 3 == 0:
            time.sleep(float(rng.uniform(0.0, 0.003)))           # one rank dawdles: the others' finishing CTAs wait in the exchange
        r, s, c = sess.search(q)
        res.append((r.tolist(), s.tobytes(), int(c)))
    # pipelined device-resident submissions
    dq = torch.from_numpy(qs).cuda()
    d_rows = torch.full((len(qs), k), -1, dtype=torch.int64, device="cuda"); d_scores = torch.zeros((len(qs), k), dtype=torch.float32, device="cuda")
    d_counts = torch.zeros((len(qs),), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    t = 0
    for i in range(len(qs)):
        t = sess.submit_device(dq[i].data_ptr(), d_rows[i].data_ptr(), d_scores[i].data_ptr(), d_counts[i].data_ptr())
    sess.wait(t)
    st = sess.stats()
    sess.close()
    r1, s1, _ = ix.search(qs[1], k)                              # and the launch path again after the session
    out[rank] = (res, d_rows.cpu().numpy().tolist(), d_scores.cpu().numpy().tobytes(), r0[0].tolist(), r1[0].tolist(), st["launches"])
    ix.close()


@pytest.mark.skipif(_ngpus() < 2, reason="needs >= 2 GPUs")
def test_resident_sessions_on_a_sharded_index(oracle):
    import torch.multiprocessing as mp
    world = min(_ngpus(), 4)
    rng = np.random.default_rng(91)
    n, d, k, nq = 30_000 * world + 17, 128, 10, 24
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[n - 5] = rows[11]                                       # a tie across shards: the lower global row wins
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    qs[2] = rows[11]
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict(); uid_q = ctx.Queue()
        procs = [ctx.Process(target=_worker_session, args=(r, world, uid_q, rows, qs, k, out, world - 1)) for r in range(world)]
        [p.start() for p in procs]
        for p in procs:
            p.join(300)
            assert p.exitcode == 0, "a rank failed or hung"
        want = [oracle.parallel_top_k_search(q, rows, k) for q in qs]
        for r in range(world):
            res, drows, dscores, r0, r1, launches = out[r]
            ds = np.frombuffer(dscores, np.float32).reshape(nq, k)
            assert launches >= 1
            assert r0 == want[0][0].tolist() and r1 == want[1][0].tolist()
            for i in range(nq):
                wi, ws = want[i]
                assert res[i][0] == wi.tolist() and res[i][1] == ws.tobytes() and res[i][2] == k, (r, i)
                assert drows[i] == wi.tolist() and ds[i].tobytes() == ws.tobytes(), (r, i)
