"""K2 (tcgen05 batched path) parity: same bar as the exact path — indices bit-exact, scores bit-identical to the
oracle (rows widened f16 -> f32) — plus evidence that the tensor cores really carried the load (no silent
exact-kernel fallback on well-separated data)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(cg, oracle, n, d, nq, k, seed, expect_no_fallback=True, rows=None, queries=None, opts=None, f32=False):
    rng = np.random.default_rng(seed)
    if rows is None:
        rows = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    if queries is None:
        queries = rng.standard_normal((nq, d)).astype(np.float32)
    ref_rows = rows if f32 else rows.astype(np.float16).astype(np.float32)
    ix = cg.Index(d, cg.F32 if f32 else cg.F16)
    try:
        ix.add(rows)
        for key, val in (opts or {}).items():
            ix.set_option(key, val)
        r, s, c = ix.search(queries, k, cg.COSINE, path=cg.PATH_TENSOR)
        st = ix.stats()
        assert st.tc_batches >= 1
        for qi in range(queries.shape[0]):
            wi, ws = oracle.parallel_top_k_search(queries[qi], ref_rows, k)
            assert int(c[qi]) == len(wi)
            assert r[qi, :len(wi)].tolist() == wi.tolist(), (qi, r[qi, :len(wi)][:8], wi[:8])
            assert np.array_equal(s[qi, :len(wi)], ws), (qi,)
        if expect_no_fallback:
            assert st.tc_fallbacks == 0, f"{st.tc_fallbacks} queries fell back to the exact kernel"
        return st
    finally:
        ix.close()


@pytest.mark.parametrize("d", [64, 128, 768, 1024])
def test_tensor_path_matches_oracle(cg, oracle, d):
    _run(cg, oracle, 60_000, d, 64, 10, seed=d)


def test_tensor_path_k100_config4_shape(cg, oracle):
    """BASELINE config 4's per-GPU shape at reduced N: d=1024 f16, batch-64, top-100."""
    _run(cg, oracle, 120_000, 1024, 64, 100, seed=4)


@pytest.mark.parametrize("nq", [1, 15, 16, 17, 100, 130])
def test_tensor_path_batch_padding_and_splitting(cg, oracle, nq):
    _run(cg, oracle, 40_000, 256, nq, 10, seed=nq)


@pytest.mark.parametrize("d", [40, 100, 1000])
def test_tensor_path_ragged_dimensions(cg, oracle, d):
    """d % 64 != 0: the TMA boxes read zeros past the row end; d % 8 != 0 exercises the exact re-score's scalar tail."""
    _run(cg, oracle, 30_000, d, 32, 10, seed=d)


def test_tensor_path_small_index_and_row_tail(cg, oracle):
    _run(cg, oracle, 8_192 + 77, 128, 16, 10, seed=1)
    _run(cg, oracle, 300, 128, 16, 10, seed=2)
    _run(cg, oracle, 5, 128, 16, 10, seed=3)


def test_tensor_path_duplicates_fall_back_but_stay_exact(cg, oracle):
    """Massive exact ties defeat the separation proof (and overflow the candidate lists): those queries must be
    re-run on the exact-order kernel and still return the lowest row indices among equals."""
    rng = np.random.default_rng(7)
    base = (rng.standard_normal((3, 128)) / 11).astype(np.float32)
    rows = np.repeat(base, 20_000, axis=0)[rng.permutation(60_000)]
    qs = np.concatenate([base[:1], rng.standard_normal((15, 128)).astype(np.float32)])
    st = _run(cg, oracle, 0, 128, 0, 20, seed=0, expect_no_fallback=False, rows=rows, queries=qs)
    assert st.tc_fallbacks >= 1


def test_auto_path_picks_tensor_for_large_f16_batches(cg, oracle):
    rng = np.random.default_rng(9)
    rows = (rng.standard_normal((50_000, 128)) / 11).astype(np.float32)
    qs = rng.standard_normal((32, 128)).astype(np.float32)
    ix = cg.Index(128, cg.F16)
    ix.add(rows)
    r, s, c = ix.search(qs, 10)                       # PATH_AUTO
    assert ix.stats().tc_batches >= 1
    r1, s1, c1 = ix.search(qs[:2], 10)                # small batch -> exact kernel
    assert ix.stats().tc_batches == 1
    assert r1.tolist() == r[:2].tolist() and s1.tobytes() == s[:2].tobytes()
    ix.close()
    # the tensor path serves cosine only
    ix = cg.Index(128, cg.F16)
    ix.add(rows)
    with pytest.raises(cg.CgvecError) as e:
        ix.search(qs, 10, cg.L2, path=cg.PATH_TENSOR)
    assert e.value.code == cg.ERR_UNSUPPORTED
    ix.close()


@pytest.mark.parametrize("d,nq,k", [(128, 16, 10), (128, 64, 10), (768, 256, 100), (1024, 64, 100), (1000, 200, 10), (256, 250, 5)])
def test_paired_cta_kernel_matches_oracle(cg, oracle, d, nq, k):
    """K2b (cta_group::2): CTA pairs, M = 256, query block streamed — up to 256 queries in one HBM pass."""
    st = _run(cg, oracle, 70_000, d, nq, k, seed=d + nq, opts={"tc_kernel": 2})
    assert st.tc_batches == 1


def test_paired_cta_kernel_row_tails_and_small_indexes(cg, oracle):
    for n in (5, 255, 256, 257, 8_192 + 300):
        _run(cg, oracle, n, 128, 32, 10, seed=n, opts={"tc_kernel": 2})


def test_auto_uses_pairs_for_batches_beyond_the_resident_limit(cg, oracle):
    st = _run(cg, oracle, 50_000, 768, 256, 10, seed=5)          # 256 queries at d=768 cannot stay resident
    assert st.tc_batches == 1


@pytest.mark.parametrize("d,nq,k,kernel", [(128, 64, 10, 1), (384, 64, 10, 1), (384, 256, 10, 2), (768, 32, 100, 1), (100, 48, 10, 2)])
def test_tf32_tensor_path_for_f32_storage(cg, oracle, d, nq, k, kernel):
    """f32 rows go through kind::tf32 (both operands lose 13 mantissa bits inside the tensor core); the exact re-score
    and the wider error bound still deliver bit-exact results (BASELINE config 5's storage type)."""
    _run(cg, oracle, 60_000, d, nq, k, seed=d * 3 + nq, f32=True, opts={"tc_kernel": kernel})


def test_symbol_resolver_dense_pass(cg, oracle):
    """U x S cosine arg-max with threshold 0.75 (codegraph-mcp/src/indexer.rs:2827-2843) as dense tensor-core
    passes (k = 1) with the winners re-scored in the reference's sequential cosine: U = 4096 references, S = 200k symbols."""
    rng = np.random.default_rng(2843)
    S, U, d = 200_000, 4096, 384
    syms = rng.standard_normal((S, d)).astype(np.float32)
    syms /= np.linalg.norm(syms, axis=1, keepdims=True)
    targets = rng.standard_normal((U, d)).astype(np.float32)
    planted = rng.choice(U, U // 2, replace=False)                       # half the references really are a known symbol + noise
    which = rng.integers(0, S, len(planted))
    targets[planted] = syms[which] + rng.normal(0, 0.3 / np.sqrt(d), (len(planted), d)).astype(np.float32)
    ix = cg.Index(d, cg.F32)
    try:
        ix.add(syms)
        got = cg.resolve_symbols(ix, targets, 0.75)
        st = ix.stats()
        assert st.tc_batches >= U // 256, "the dense path should have carried the batch"
        assert st.tc_fallbacks <= U // 50
        hit = {int(u): int(w) for u, w in zip(planted, which)}
        check = list(planted[:40]) + [u for u in range(U) if u not in hit][:24]
        for u in check:
            wi, ws = oracle.parallel_top_k_search(targets[u], syms, 1, form=oracle.FORM_SEQ)
            want = (int(wi[0]), float(ws[0])) if ws[0] > 0.75 else None
            assert got[u] == want, (u, got[u], want)
        assert all(got[u] is not None and got[u][0] == hit[u] for u in list(hit)[:500])
        assert sum(g is None for g in got) >= U // 2 - 8               # unrelated references stay unresolved
    finally:
        ix.close()
