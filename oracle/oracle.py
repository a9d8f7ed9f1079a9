"""ctypes loader for the CPU oracle (oracle/cgvec_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under codegraph-rust_b200/ may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcgvec_oracle.so")

COSINE, DOT, L2 = 0, 1, 2
FORM_ADAPTIVE, FORM_SCALAR, FORM_SEQ = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cgvec_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp, u64p, u8p, u16p = C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_uint16)
        for name in ("cg_cosine_similarity_avx2", "cg_cosine_similarity_avx2_emul", "cg_dot_product_avx2",
                     "cg_l2_distance_avx2", "cg_cosine_similarity_scalar", "cg_adaptive_cosine_similarity",
                     "cg_cosine_similarity_seq", "cg_cosine_distance_seq"):
            f = getattr(L, name); f.restype = C.c_float; f.argtypes = [fp, fp, C.c_size_t]
        L.cg_normalize_avx2.restype = None; L.cg_normalize_avx2.argtypes = [fp, C.c_size_t]
        L.cg_scores.restype = None
        L.cg_scores.argtypes = [fp, fp, C.c_uint64, C.c_size_t, C.c_size_t, C.c_int, C.c_int, fp]
        L.cg_parallel_top_k_search.restype = C.c_uint64
        L.cg_parallel_top_k_search.argtypes = [fp, fp, C.c_uint64, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.c_int, u64p, fp]
        for name in ("cg_search_baseline", "cg_inmemory_search_similar"):
            f = getattr(L, name); f.restype = C.c_uint64
            f.argtypes = [fp, fp, C.c_uint64, C.c_size_t, C.c_size_t, C.c_uint64, u64p, fp]
        L.cg_compute_distances_cpu.restype = C.c_uint64
        L.cg_compute_distances_cpu.argtypes = [fp, fp, C.c_uint64, C.c_size_t, C.c_uint64, fp]
        L.cg_normalize_scores.restype = None; L.cg_normalize_scores.argtypes = [fp, C.c_size_t]
        L.cg_prefetch_k_basic.restype = C.c_uint64; L.cg_prefetch_k_basic.argtypes = [C.c_uint64]
        L.cg_prefetch_k_filtered.restype = C.c_uint64; L.cg_prefetch_k_filtered.argtypes = [C.c_uint64]
        L.cg_quantize_unit_i8.restype = C.c_int8; L.cg_quantize_unit_i8.argtypes = [C.c_float]
        L.cg_quantize_batch_u8.restype = None; L.cg_quantize_batch_u8.argtypes = [fp, C.c_uint64, C.c_size_t, u8p]
        L.cg_search_optimized_i8.restype = C.c_uint64
        L.cg_search_optimized_i8.argtypes = [fp, C.c_size_t, u8p, C.c_uint64, C.c_size_t, C.c_uint64, u64p, fp]
        L.cg_siphash13_2xu64.restype = C.c_uint64; L.cg_siphash13_2xu64.argtypes = [C.c_uint64, C.c_uint64]
        L.cg_generate_optimization_vectors.restype = None
        L.cg_generate_optimization_vectors.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, fp]
        L.cg_hash_text_embedding.restype = None
        L.cg_hash_text_embedding.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, fp]
        L.cg_widen_f16.restype = None; L.cg_widen_f16.argtypes = [u16p, C.c_uint64, fp]
        L.cg_narrow_f16.restype = None; L.cg_narrow_f16.argtypes = [fp, C.c_uint64, u16p]
        L.cg_synth_rows.restype = None
        L.cg_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, C.c_int, C.c_int, fp]
        L.cg_vecs_create.restype = C.c_void_p; L.cg_vecs_create.argtypes = [fp, C.c_uint64, C.c_size_t]
        L.cg_vecs_destroy.restype = None; L.cg_vecs_destroy.argtypes = [C.c_void_p]
        L.cg_max_threads.restype = C.c_int
        L.cg_affinity_threads.restype = C.c_int
        L.cg_fair_top_k_search_multi.restype = None
        L.cg_fair_top_k_search_multi.argtypes = [fp, C.c_uint64, fp, C.c_uint64, C.c_size_t, C.c_uint64, C.c_int, u64p, fp, u64p]
        L.cg_have_avx2.restype = C.c_int
        L.cg_parallel_top_k_search_mt.restype = C.c_uint64
        L.cg_parallel_top_k_search_mt.argtypes = [fp, C.c_void_p, C.c_uint64, C.c_int, u64p, fp]
        L.cg_fair_top_k_search_mt.restype = C.c_uint64
        L.cg_fair_top_k_search_mt.argtypes = [fp, fp, C.c_uint64, C.c_size_t, C.c_uint64, C.c_int, u64p, fp]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pair_fn(name):
    def f(a, b):
        a, b = _f32(a), _f32(b)
        assert a.shape == b.shape and a.ndim == 1
        return float(getattr(lib(), name)(_fp(a), _fp(b), a.size))
    f.__name__ = name
    return f


cosine_similarity_avx2 = _pair_fn("cg_cosine_similarity_avx2")            # simd_ops.rs:15-78
cosine_similarity_avx2_emul = _pair_fn("cg_cosine_similarity_avx2_emul")
dot_product_avx2 = _pair_fn("cg_dot_product_avx2")                        # simd_ops.rs:149-183
l2_distance_avx2 = _pair_fn("cg_l2_distance_avx2")                        # simd_ops.rs:105-143
cosine_similarity_scalar = _pair_fn("cg_cosine_similarity_scalar")        # simd_ops.rs:257-278
adaptive_cosine_similarity = _pair_fn("cg_adaptive_cosine_similarity")    # simd_ops.rs:281-295
cosine_similarity_seq = _pair_fn("cg_cosine_similarity_seq")              # search.rs:519-533
cosine_distance_seq = _pair_fn("cg_cosine_distance_seq")                  # optimization.rs:404-418


def normalize_avx2(v):
    v = _f32(v).copy()
    lib().cg_normalize_avx2(_fp(v), v.size)
    return v


def scores(query, rows, metric=COSINE, form=FORM_ADAPTIVE):
    q, r = _f32(query), _f32(rows)
    out = np.empty(r.shape[0], np.float32)
    lib().cg_scores(_fp(q), _fp(r), r.shape[0], r.shape[1], r.shape[1], metric, form, _fp(out))
    return out


def parallel_top_k_search(query, rows, k, metric=COSINE, form=FORM_ADAPTIVE):
    """simd_ops.rs:361-383 -> (indices u64[m], scores f32[m]), m = min(k, n)."""
    q, r = _f32(query), _f32(rows)
    n, d = r.shape
    m = min(int(k), n)
    idx = np.empty(max(m, 1), np.uint64); sc = np.empty(max(m, 1), np.float32)
    got = lib().cg_parallel_top_k_search(_fp(q), _fp(r), n, d, d, k, metric, form,
                                         idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc))
    return idx[:got].copy(), sc[:got].copy()


def _limit_fn(name):
    def f(query, rows, limit):
        q, r = _f32(query), _f32(rows)
        n, d = r.shape
        m = min(int(limit), n)
        idx = np.empty(max(m, 1), np.uint64); sc = np.empty(max(m, 1), np.float32)
        got = getattr(lib(), name)(_fp(q), _fp(r), n, d, d, limit,
                                   idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc))
        return idx[:got].copy(), sc[:got].copy()
    return f


search_baseline = _limit_fn("cg_search_baseline")                    # optimization.rs:376-402
inmemory_search_similar = _limit_fn("cg_inmemory_search_similar")    # core integration/graph_vector.rs:479-494


def compute_distances_cpu(query, flat_rows, dimension, limit):          # gpu.rs:297-322
    q, r = _f32(query), _f32(flat_rows).reshape(-1)
    assert r.size % dimension == 0
    n = r.size // dimension
    out = np.empty(max(min(limit, n), 1), np.float32)
    got = lib().cg_compute_distances_cpu(_fp(q), _fp(r), n, dimension, limit, _fp(out))
    return out[:got].copy()


def normalize_scores(s):                                                # search.rs:574-592
    s = _f32(s).copy()
    lib().cg_normalize_scores(_fp(s), s.size)
    return s


def prefetch_k_basic(limit):       # search.rs:113
    return int(lib().cg_prefetch_k_basic(limit))


def prefetch_k_filtered(limit):    # search.rs:276
    return int(lib().cg_prefetch_k_filtered(limit))


def search_by_embedding(query, rows, limit):
    """search.rs:91-144 SemanticSearch::search_by_embedding over an exact store:
    over-fetch prefetch_k ids through the trait (graph_vector.rs:479-494 semantics), rescore each
    with search.rs:519-533, stable sort desc, truncate, min-max normalise.  -> (idx, norm_scores, raw)"""
    pk = prefetch_k_basic(limit)
    ids, _ = inmemory_search_similar(query, rows, pk)
    q = _f32(query); r = _f32(rows)
    raw = np.array([cosine_similarity_seq(q, r[int(i)]) for i in ids], np.float32)
    order = sorted(range(len(ids)), key=lambda j: (-raw[j], j)) if not np.isnan(raw).any() else list(range(len(ids)))
    order = order[:limit]
    ids2 = ids[order]; raw2 = raw[order]
    return ids2, normalize_scores(raw2), raw2


def quantize_batch_u8(rows):                                            # optimization.rs:212-224,268-274
    r = _f32(rows)
    out = np.empty(r.shape, np.uint8)
    lib().cg_quantize_batch_u8(_fp(r), r.shape[0], r.shape[1], out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def search_optimized_i8(query, codes, limit):                           # optimization.rs:63-150
    q = _f32(query); c = np.ascontiguousarray(codes, np.uint8)
    n, d = c.shape
    cap = max(min(max(limit, 1), n), 1)
    idx = np.empty(cap, np.uint64); sc = np.empty(cap, np.float32)
    got = lib().cg_search_optimized_i8(_fp(q), q.size, c.ctypes.data_as(C.POINTER(C.c_uint8)), n, d, limit,
                                       idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc))
    return idx[:got].copy(), sc[:got].copy()


def siphash13(m0, m1):
    return int(lib().cg_siphash13_2xu64(m0 & (2**64 - 1), m1 & (2**64 - 1)))


def generate_optimization_vectors(count, dim, seed):                    # model_optimization_tests.rs:36-58
    out = np.empty((count, dim), np.float32)
    lib().cg_generate_optimization_vectors(count, dim, seed, _fp(out))
    return out


def hash_text_embedding(text: str, dimension: int = 384):               # search.rs:178-205
    b = text.encode("utf-8")
    out = np.empty(dimension, np.float32)
    lib().cg_hash_text_embedding(b, len(b), dimension, _fp(out))
    return out


def widen_f16(h):
    h = np.ascontiguousarray(h, np.uint16)
    out = np.empty(h.shape, np.float32)
    lib().cg_widen_f16(h.ctypes.data_as(C.POINTER(C.c_uint16)), h.size, _fp(out))
    return out


def narrow_f16(f):
    f = _f32(f)
    out = np.empty(f.shape, np.uint16)
    lib().cg_narrow_f16(_fp(f), f.size, out.ctypes.data_as(C.POINTER(C.c_uint16)))
    return out


class RefVecs:
    """&[Vec<f32>] — N separate heap rows, as parallel_top_k_search receives them (simd_ops.rs:363)."""

    def __init__(self, rows):
        r = _f32(rows)
        self.n, self.d = r.shape
        self.h = lib().cg_vecs_create(_fp(r), self.n, self.d)

    def top_k_mt(self, query, k, threads=0):
        q = _f32(query)
        m = max(min(k, self.n), 1)
        idx = np.empty(m, np.uint64); sc = np.empty(m, np.float32)
        got = lib().cg_parallel_top_k_search_mt(_fp(q), self.h, k, threads,
                                                idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc))
        return idx[:got].copy(), sc[:got].copy()

    def close(self):
        if self.h:
            lib().cg_vecs_destroy(self.h); self.h = None

    __del__ = close


def fair_top_k_mt(query, rows, k, threads=0):
    q, r = _f32(query), _f32(rows)
    m = max(min(k, r.shape[0]), 1)
    idx = np.empty(m, np.uint64); sc = np.empty(m, np.float32)
    got = lib().cg_fair_top_k_search_mt(_fp(q), _fp(r), r.shape[0], r.shape[1], k, threads,
                                        idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc))
    return idx[:got].copy(), sc[:got].copy()


def fair_top_k_multi(queries, rows, k, threads=0):
    """cg_parallel_top_k_search's outputs for several queries in one pass over `rows` -> list of (indices, scores)."""
    qs, r = _f32(queries), _f32(rows)
    nq = qs.shape[0]
    idx = np.zeros((nq, max(k, 1)), np.uint64); sc = np.zeros((nq, max(k, 1)), np.float32); cnt = np.zeros(nq, np.uint64)
    lib().cg_fair_top_k_search_multi(_fp(qs), nq, _fp(r), r.shape[0], r.shape[1], k, threads,
                                     idx.ctypes.data_as(C.POINTER(C.c_uint64)), _fp(sc), cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    return [(idx[i, :int(cnt[i])].copy(), sc[i, :int(cnt[i])].copy()) for i in range(nq)]


def merge_top_k(parts, k):
    """Merges per-chunk (global indices, scores) lists under the result contract (best first, ties -> lower row, NaN last)."""
    idx = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, np.uint64)
    sc = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.float32)
    nan = np.isnan(sc)
    key = np.where(nan, -np.inf, sc.astype(np.float64))
    order = np.lexsort((idx, -key, nan))            # primary: non-NaN first, then score descending, then row ascending
    order = order[:k]
    return idx[order], sc[order]


def max_threads():
    return int(lib().cg_max_threads())


def affinity_threads():
    return int(lib().cg_affinity_threads())


def synth_rows(seed, first_row, n, d, unit_norm=True, f16=False, threads=0):
    """Host twin of cgvec_fill_synthetic (rows first_row .. first_row+n)."""
    out = np.empty((n, d), np.float32)
    lib().cg_synth_rows(seed, first_row, n, d, 1 if unit_norm else 0, 1 if f16 else 0, threads, _fp(out))
    return out
