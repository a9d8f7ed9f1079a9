"""CPU-side checks (no GPU needed): the C-ABI library loads and exports every symbol include/cgvec.h
declares, fails loudly (no CPU fallback) without a device, and its host-side helpers agree with the oracle."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol(cg):
    lib = cg.load_library()
    hdr = open(os.path.join(ROOT, "include", "cgvec.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cgvec_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(cg.EXPORTS), declared ^ set(cg.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", cg.lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cgvec_[a-z0-9_]+)", out))
    assert declared <= exported
    # nothing but the ABI leaks out of the shared object
    leaked = [l for l in out.splitlines() if " T " in l and "cgvec_" not in l]
    assert not leaked, leaked[:5]


def test_library_is_sm100a_native(cg):
    """The scan kernel must carry the TMA-engine bulk copy (SASS UBLKCP) and mbarrier ops — not a generic build."""
    out = subprocess.run(["cuobjdump", "-sass", cg.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    assert "UBLKCP" in out.stdout and "SYNCS" in out.stdout


@pytest.mark.skipif(_has_gpu(), reason="this container check only applies without a GPU")
def test_no_cpu_fallback_without_device(cg):
    with pytest.raises(cg.CgvecError) as ei:
        cg.Index(768)
    assert ei.value.code == cg.ERR_NO_DEVICE
    assert "no CPU fallback" in ei.value.msg
    with pytest.raises(cg.CgvecError):
        cg.ParallelVectorOps.parallel_top_k_search(np.ones(8, np.float32), np.ones((4, 8), np.float32), 2)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "codegraph-rust_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h", ".hpp", ".cpp", ".rs")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "cgvec_oracle" not in txt, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libcgvec_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_argument_validation_never_aborts(cg):
    import ctypes as C
    lib = cg.load_library()
    h = C.c_void_p()
    assert lib.cgvec_create(0, 0, None, 1, C.byref(h)) == cg.ERR_BAD_DIM
    assert lib.cgvec_create(8, 7, None, 1, C.byref(h)) == cg.ERR_BAD_ARG
    assert lib.cgvec_create(8, 0, None, 0, C.byref(h)) == cg.ERR_BAD_ARG
    assert lib.cgvec_create(8, 0, None, 1, None) == cg.ERR_BAD_ARG
    assert lib.cgvec_add(None, None, None, 5) == cg.ERR_BAD_ARG
    assert lib.cgvec_search(None, None, 1, 1, 0, None, None, None, None) == cg.ERR_BAD_ARG
    assert lib.cgvec_destroy(None) == 0
    assert lib.cgvec_len(None) == 0
    assert lib.cgvec_create_rank(8, 0, 0, 2, 2, None, 0, C.byref(h)) == cg.ERR_BAD_ARG
    assert b"rank" in lib.cgvec_last_error()
    # resident sessions and streams: NULL handles are errors (or no-ops for close), never a crash
    t = C.c_uint32(); ms = C.c_float()
    assert lib.cgvec_serve_open(None, 10, 0, C.byref(h)) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_search(None, None, None, None, None, None) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_submit(None, None, 0, None, None, None, C.byref(t)) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_wait(None, 1) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_pause(None) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_timer_start(None) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_timer_stop(None, C.byref(ms)) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_set(None, b"idle_us", 1) == cg.ERR_BAD_ARG
    assert lib.cgvec_serve_close(None) == 0
    assert lib.cgvec_stream_close(None) == 0


def test_shard_range_covers_and_partitions(cg):
    for n in (0, 1, 7, 8, 9, 1000, 10**6 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [cg.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            per = -(-n // world)
            assert all(e - b <= per for b, e in spans)


def test_prefetch_and_normalise_match_oracle(cg, oracle):
    for lim in (0, 1, 3, 5, 10, 100, 1000):
        assert cg.prefetch_k_basic(lim) == oracle.prefetch_k_basic(lim)
        assert cg.prefetch_k_filtered(lim) == oracle.prefetch_k_filtered(lim)
    rng = np.random.default_rng(0)
    for n in (1, 2, 17):
        s = rng.standard_normal(n).astype(np.float32)
        assert cg.normalize_scores(s).tobytes() == oracle.normalize_scores(s).tobytes()
    assert cg.normalize_scores(np.float32([0.25, 0.25])).tolist() == [0.0, 0.0]


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_merge_topk_host_equals_global_topk(cg, oracle, metric):
    """Top-k of a union == merge of per-shard top-k, incl. ties across shards and short shards."""
    rng = np.random.default_rng(4)
    n, d, k, world = 1000, 48, 20, 3
    rows = rng.standard_normal((n, d)).astype(np.float32)
    rows[500] = rows[10]; rows[900] = rows[10]          # exact ties across shards
    q = rows[10] + 0.01 * rng.standard_normal(d).astype(np.float32)
    m = oracle.COSINE if metric == "cosine" else oracle.L2
    want_idx, want_sc = oracle.parallel_top_k_search(q, rows, k, metric=m)
    prow = np.zeros((world, k), np.uint64); psc = np.zeros((world, k), np.float32); pcnt = np.zeros(world, np.uint32)
    for r in range(world):
        b, e = cg.shard_range(n, world, r)
        idx, sc = oracle.parallel_top_k_search(q, rows[b:e], k, metric=m)
        prow[r, :len(idx)] = idx + b; psc[r, :len(idx)] = sc; pcnt[r] = len(idx)
    got_idx, got_sc = cg.merge_topk_host(prow, psc, pcnt, k, ascending=(metric == "l2"))
    assert got_idx.tolist() == want_idx.tolist()
    assert got_sc.tobytes() == want_sc.tobytes()
    # a shard with fewer than k rows
    pcnt2 = pcnt.copy(); pcnt2[2] = 3
    got2, _ = cg.merge_topk_host(prow, psc, pcnt2, k, ascending=(metric == "l2"))
    assert len(got2) == k


def test_synth_mirror_properties():
    """tests/synth.py is the host twin of the device generator: check its own invariants here (the GPU
    test checks device == mirror bit for bit)."""
    from tests import synth
    x = synth.synth_rows(0xC0DE6A9F, np.arange(64), 768, unit_norm=True)
    assert x.dtype == np.float32 and x.shape == (64, 768)
    n = np.linalg.norm(x.astype(np.float64), axis=1)
    assert np.all(np.abs(n - 1) < 1e-6)
    raw = synth.synth_raw(1, np.arange(2000), 64)
    assert abs(float(raw.mean())) < 0.01 and 0.5 < float(raw.std()) < 0.65
    assert synth.synth_raw(1, [5], 8).tobytes() == synth.synth_raw(1, [5], 8).tobytes()
    assert synth.synth_raw(1, [5], 8).tobytes() != synth.synth_raw(2, [5], 8).tobytes()


def test_multi_device_placement_is_a_bijection(cg):
    """Single-process multi-device index: global rows <-> (device, local row) through 1024-row blocks dealt round-robin."""
    for G in (1, 2, 3, 8):
        for n in (0, 1, 1023, 1024, 1025, 5000, 8 * 1024 * 3 + 17):
            counts = [cg.multi_local_count(G, s, n) for s in range(G)]
            assert sum(counts) == n
            assert max(counts) - min(counts) <= 1024
            seen = set()
            for g in list(range(min(n, 3000))) + list(range(max(0, n - 3000), n)):
                s, l = cg.multi_locate(G, g)
                assert 0 <= s < G and l < counts[s]
                seen.add((s, l))
                if g + 1 < n:                       # order inside a device follows global order (tie rule survives)
                    s2, l2 = cg.multi_locate(G, g + 1)
                    assert (s2 != s) or (l2 == l + 1)
            assert len(seen) == len(set(list(range(min(n, 3000))) + list(range(max(0, n - 3000), n))))


def test_merge_topk_host_contract_with_ties_nan_and_short_lists(cg):
    """Property check of cgvec_merge_topk_host against a brute-force sort under the result contract (best first, ties ->
    lower row, NaN last, -0.0 == +0.0), for random partitions, duplicate scores, NaNs and lists shorter than k."""
    rng = np.random.default_rng(123)
    for trial in range(60):
        parts = int(rng.integers(1, 7)); k = int(rng.integers(1, 12)); asc = bool(rng.integers(0, 2))
        pool_scores = rng.choice(np.float32([0.0, -0.0, 0.25, 0.25, 0.5, -1.0, np.nan, 1.0, np.inf, -np.inf]), size=parts * k)
        pool_scores = np.where(rng.random(parts * k) < 0.5, pool_scores, rng.standard_normal(parts * k).astype(np.float32)).astype(np.float32)
        pool_rows = rng.permutation(parts * k * 3)[: parts * k].astype(np.uint64)

        def key(i):
            s = pool_scores[i]
            if np.isnan(s):
                return (1, 0.0, int(pool_rows[i]))
            v = float(s) + 0.0
            return (0, v if asc else -v, int(pool_rows[i]))

        rows = np.zeros((parts, k), np.uint64); scores = np.zeros((parts, k), np.float32); counts = np.zeros(parts, np.uint32)
        members = []
        for p in range(parts):
            idx = list(range(p * k, (p + 1) * k))
            cnt = int(rng.integers(0, k + 1))
            idx = sorted(idx, key=key)[:cnt]                       # each partial list is itself sorted under the contract
            counts[p] = cnt
            for j, i in enumerate(idx):
                rows[p, j] = pool_rows[i]; scores[p, j] = pool_scores[i]
            members += idx
        want = sorted(members, key=key)[:k]
        got_rows, got_scores = cg.merge_topk_host(rows, scores, counts, k, ascending=asc)
        assert got_rows.tolist() == [int(pool_rows[i]) for i in want], (trial, asc)
        assert np.array_equal(got_scores, np.float32([pool_scores[i] for i in want]), equal_nan=True)


def test_enable_gpu_switch_and_device_list(cg, monkeypatch):
    """cgvec_create_from_env: PerformanceConfig.enable_gpu (config_manager.rs:362-364, default false :419) with the CODEGRAPH_ENABLE_GPU
    override, and CODEGRAPH_B200_DEVICES parsing.  Off -> ERR_DISABLED (host keeps its CPU store); malformed lists never abort."""
    monkeypatch.delenv("CODEGRAPH_ENABLE_GPU", raising=False)
    monkeypatch.delenv("CODEGRAPH_B200_DEVICES", raising=False)
    with pytest.raises(cg.CgvecError) as e:
        cg.Index.from_env(64, enable_gpu=False)
    assert e.value.code == cg.ERR_DISABLED
    monkeypatch.setenv("CODEGRAPH_ENABLE_GPU", "false")
    with pytest.raises(cg.CgvecError) as e:
        cg.Index.from_env(64, enable_gpu=True)                 # the environment wins over the config field
    assert e.value.code == cg.ERR_DISABLED
    monkeypatch.setenv("CODEGRAPH_ENABLE_GPU", "1")
    for bad in ("0,x", "1,,2", "-1", "0", "99999"):
        monkeypatch.setenv("CODEGRAPH_B200_DEVICES", bad)
        with pytest.raises(cg.CgvecError) as e:
            cg.Index.from_env(64)
        assert e.value.code in (cg.ERR_BAD_ARG, cg.ERR_NO_DEVICE, cg.ERR_UNSUPPORTED), bad
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        for ok in ("0,1", "2", "all"):
            monkeypatch.setenv("CODEGRAPH_B200_DEVICES", ok)
            with pytest.raises(cg.CgvecError) as e:
                cg.Index.from_env(64)
            assert e.value.code == cg.ERR_NO_DEVICE


def test_auto_cost_model_agrees_with_the_committed_measurements(cg):
    """CGVEC_PATH_AUTO's cost model (host_tensor.inl) against the B200 measurements it was fitted to
    (profiles/r02_exact_vs_tensor_small_batches.txt, tools/bench_paths.py): it must pick the measured-faster family wherever the
    two differ by more than 15 %, and its estimates must stay within a factor 1.5 of the measured call times."""
    import json
    path = os.path.join(ROOT, "profiles", "r02_exact_vs_tensor_small_batches.txt")
    checked = picked = 0
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        r = json.loads(line)
        if r["nq"] < 2 or r.get("exact_ms") is None or r.get("tensor_ms") is None:
            continue                                            # batch-1 stays on the exact-order kernel by rule, not by cost
        dt = cg.F32 if r["dtype"] == "f32" else cg.F16
        te, tt = cg.path_cost_model(dt, 768, r["rows"], r["nq"], 128)
        assert r["exact_ms"] / 1.5 <= te <= r["exact_ms"] * 1.5, (r, te)
        assert r["tensor_ms"] / 1.5 <= tt <= r["tensor_ms"] * 1.5, (r, tt)
        checked += 1
        if max(r["exact_ms"], r["tensor_ms"]) > 1.15 * min(r["exact_ms"], r["tensor_ms"]):
            assert (tt < te) == (r["tensor_ms"] < r["exact_ms"]), (r, te, tt)
            picked += 1
    assert checked >= 30 and picked >= 20
