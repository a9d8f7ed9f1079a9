"""codegraph-rust_b200 — ctypes binding over libcgvec_b200.so (include/cgvec.h) plus Python mirrors of
the reference interfaces that sit on the similarity-search path, so the parity tests read like the
reference's own tests:

    ParallelVectorOps.parallel_top_k_search      crates/codegraph-vector/src/simd_ops.rs:361-383
    B200VectorStore  (trait VectorStore)         crates/codegraph-core/src/traits.rs:11-16
    B200Backend      (trait SurrealVectorBackend) crates/codegraph-vector/src/surreal_store.rs:11-22
    SemanticSearch.search_by_embedding           crates/codegraph-vector/src/search.rs:91-144
    GpuAcceleration.upload_vectors / compute_distances   crates/codegraph-vector/src/gpu.rs:221-291

Every score and every top-k decision is computed by the CUDA library; this module only marshals
buffers.  It never imports the oracle and has no CPU scoring path: without the built extension (or
without an sm_100 GPU) it raises.
"""
from __future__ import annotations

import ctypes as C
import os
import uuid as _uuid
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _build

F32, F16 = 0, 1
COSINE, DOT, L2 = 0, 1, 2
FORMULA_SIMD, FORMULA_SCALAR, FORMULA_SEQ, FORMULA_BASELINE = 0, 1, 2, 3
PATH_AUTO, PATH_EXACT, PATH_TENSOR = 0, 1, 2

OK, ERR_BAD_ARG, ERR_BAD_DIM, ERR_OOM, ERR_CUDA, ERR_NCCL, ERR_NOT_FOUND, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_DISABLED = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9)

EXPORTS = [
    "cgvec_create", "cgvec_create_from_env", "cgvec_create_rank", "cgvec_nccl_unique_id", "cgvec_destroy", "cgvec_reserve", "cgvec_add",
    "cgvec_add_f16", "cgvec_normalize_rows", "cgvec_fill_synthetic", "cgvec_len", "cgvec_dim", "cgvec_search",
    "cgvec_search_ex", "cgvec_get", "cgvec_get_row", "cgvec_get_rows", "cgvec_row_of_id", "cgvec_rescore", "cgvec_distances_first",
    "cgvec_quantize_i8", "cgvec_get_codes_i8", "cgvec_search_i8", "cgvec_save_flat", "cgvec_load_flat", "cgvec_shard_range", "cgvec_multi_locate", "cgvec_multi_local_count", "cgvec_merge_topk_host", "cgvec_path_cost_model", "cgvec_prefetch_k_basic", "cgvec_prefetch_k_filtered",
    "cgvec_normalize_scores", "cgvec_get_stats", "cgvec_set_option", "cgvec_get_trace", "cgvec_last_error", "cgvec_version",
    "cgvec_stream_open", "cgvec_stream_submit", "cgvec_stream_flush", "cgvec_stream_close",
    "cgvec_serve_open", "cgvec_serve_submit", "cgvec_serve_wait", "cgvec_serve_search", "cgvec_serve_pause", "cgvec_serve_stats",
    "cgvec_serve_set", "cgvec_serve_close", "cgvec_serve_timer_start", "cgvec_serve_timer_stop",
]


class CgvecError(RuntimeError):
    """Maps to CodeGraphError::Vector(msg) (codegraph-core/src/error.rs:18-19)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"cgvec error {code}: {msg}")
        self.code = code
        self.msg = msg


class SearchOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("metric", C.c_int), ("formula", C.c_int), ("path", C.c_int),
                ("stream", C.c_void_p), ("device_io", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("searches", C.c_uint64), ("rows", C.c_uint64),
                ("bytes_resident", C.c_uint64), ("sm_count", C.c_uint32), ("grid", C.c_uint32), ("block", C.c_uint32),
                ("smem_bytes", C.c_uint32), ("stages", C.c_uint32), ("tile_rows", C.c_uint32),
                ("last_scan_ms", C.c_float), ("scan_ms_total", C.c_double), ("scans_timed", C.c_uint64),
                ("tc_batches", C.c_uint64), ("tc_fallbacks", C.c_uint64), ("exchange_mode", C.c_uint32), ("reserved0", C.c_uint32),
                ("tc_main_ms_total", C.c_double), ("tc_main_timed", C.c_uint64),
                ("coalesced_batches", C.c_uint64), ("coalesced_queries", C.c_uint64)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load_library(build: bool = True):
    """Loads libcgvec_b200.so (building it first if sources are newer). Fails loudly if it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if build:
        _build.build_lib()
    if not os.path.exists(_build.LIB):
        raise CgvecError(ERR_NO_DEVICE, f"{_build.LIB} is not built; run __graft_entry__.build()")
    L = C.CDLL(_build.LIB)
    vp, u64p, fp, u32p, u8p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_void_p
    L.cgvec_create.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.cgvec_create_from_env.argtypes = [C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]
    L.cgvec_create_rank.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_uint64, C.POINTER(vp)]
    L.cgvec_nccl_unique_id.argtypes = [vp]
    L.cgvec_destroy.argtypes = [vp]
    L.cgvec_reserve.argtypes = [vp, C.c_uint64]
    L.cgvec_add.argtypes = [vp, u8p, vp, C.c_uint64]
    L.cgvec_add_f16.argtypes = [vp, u8p, vp, C.c_uint64]
    L.cgvec_normalize_rows.argtypes = [vp]
    L.cgvec_fill_synthetic.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
    L.cgvec_len.argtypes = [vp]; L.cgvec_len.restype = C.c_uint64
    L.cgvec_dim.argtypes = [vp]; L.cgvec_dim.restype = C.c_uint32
    L.cgvec_search.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, vp, vp]
    L.cgvec_search_ex.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(SearchOpts), vp, vp, vp, vp]
    L.cgvec_get.argtypes = [vp, vp, vp]
    L.cgvec_get_row.argtypes = [vp, C.c_uint64, vp]
    L.cgvec_get_rows.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.cgvec_row_of_id.argtypes = [vp, vp, u64p]
    L.cgvec_rescore.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, vp]
    L.cgvec_distances_first.argtypes = [vp, vp, C.c_uint64, vp, u64p]
    L.cgvec_quantize_i8.argtypes = [vp]
    L.cgvec_get_codes_i8.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
    L.cgvec_search_i8.argtypes = [vp, vp, C.c_uint32, vp, vp, u32p]
    L.cgvec_save_flat.argtypes = [vp, C.c_char_p]
    L.cgvec_load_flat.argtypes = [vp, C.c_char_p, u64p]
    L.cgvec_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p, u64p]
    L.cgvec_multi_locate.argtypes = [C.c_uint32, C.c_uint64, u32p, u64p]
    L.cgvec_multi_local_count.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]; L.cgvec_multi_local_count.restype = C.c_uint64
    L.cgvec_merge_topk_host.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, u32p]
    L.cgvec_prefetch_k_basic.argtypes = [C.c_uint64]; L.cgvec_prefetch_k_basic.restype = C.c_uint64
    L.cgvec_prefetch_k_filtered.argtypes = [C.c_uint64]; L.cgvec_prefetch_k_filtered.restype = C.c_uint64
    L.cgvec_normalize_scores.argtypes = [vp, C.c_size_t]; L.cgvec_normalize_scores.restype = None
    L.cgvec_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.cgvec_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.cgvec_get_trace.argtypes = [vp, vp, C.c_uint32, u32p]
    L.cgvec_stream_open.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(vp)]
    L.cgvec_stream_submit.argtypes = [vp, vp, C.c_uint32, vp, vp, vp, u32p]
    L.cgvec_stream_flush.argtypes = [vp, vp, vp, vp, u32p]
    L.cgvec_stream_close.argtypes = [vp]
    L.cgvec_serve_open.argtypes = [vp, C.c_uint32, C.c_int, C.POINTER(vp)]
    L.cgvec_serve_submit.argtypes = [vp, vp, C.c_int, vp, vp, vp, u32p]
    L.cgvec_serve_wait.argtypes = [vp, C.c_uint32]
    L.cgvec_serve_search.argtypes = [vp, vp, vp, vp, vp, vp]
    L.cgvec_serve_pause.argtypes = [vp]
    L.cgvec_serve_stats.argtypes = [vp, u64p, u64p]
    L.cgvec_serve_set.argtypes = [vp, C.c_char_p, C.c_int64]
    L.cgvec_serve_close.argtypes = [vp]
    L.cgvec_serve_timer_start.argtypes = [vp]
    L.cgvec_serve_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.cgvec_last_error.restype = C.c_char_p
    L.cgvec_version.restype = C.c_char_p
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise CgvecError(rc, load_library().cgvec_last_error().decode("utf-8", "replace"))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ids_to_bytes(ids) -> Optional[np.ndarray]:
    if ids is None:
        return None
    out = np.empty((len(ids), 16), np.uint8)
    for i, x in enumerate(ids):
        b = x.bytes if isinstance(x, _uuid.UUID) else bytes(x)
        assert len(b) == 16
        out[i] = np.frombuffer(b, np.uint8)
    return out


def path_cost_model(dtype: int, dim: int, rows: int, nq: int, tensor_batch_limit: int = 128):
    """(exact_ms, tensor_ms) of AUTO's cost model for one call with nq queries."""
    L = load_library()
    L.cgvec_path_cost_model.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    a, b = C.c_double(), C.c_double()
    _check(L.cgvec_path_cost_model(dtype, dim, rows, nq, tensor_batch_limit, C.byref(a), C.byref(b)))
    return float(a.value), float(b.value)


def version() -> str:
    return load_library().cgvec_version().decode()


# -------------------------------------------------------------------------------------------------
# thin object wrappers over the C ABI
# -------------------------------------------------------------------------------------------------
class QueryStream:
    """cgvec_stream_*: submit(batch i+1) uploads it behind the search of batch i and returns batch i's results
    (None on the first call); flush() drains the last batch.  Results: (rows u64[nq,k], scores f32[nq,k], counts u32[nq])."""

    def __init__(self, index: "Index", max_batch: int, k: int, metric: int = COSINE, path: int = PATH_AUTO):
        self._ix, self.max_batch, self.k = index, max_batch, k
        self._h = C.c_void_p()
        _check(load_library().cgvec_stream_open(index._h, max_batch, k, metric, path, C.byref(self._h)))
        self._rows = np.empty((max_batch, k), np.uint64); self._scores = np.empty((max_batch, k), np.float32)
        self._counts = np.empty(max_batch, np.uint32); self._nq = C.c_uint32()

    def _result(self):
        n = int(self._nq.value)
        if n == 0:
            return None
        return self._rows[:n].copy(), self._scores[:n].copy(), self._counts[:n].copy()

    def submit(self, queries: np.ndarray):
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim != 2 or q.shape[1] != self._ix.dim:
            raise CgvecError(ERR_BAD_DIM, f"queries have shape {q.shape}, index dimension is {self._ix.dim}")
        _check(load_library().cgvec_stream_submit(self._h, _ptr(q), q.shape[0], _ptr(self._rows), _ptr(self._scores), _ptr(self._counts),
                                                  C.byref(self._nq)))
        return self._result()

    def flush(self):
        _check(load_library().cgvec_stream_flush(self._h, _ptr(self._rows), _ptr(self._scores), _ptr(self._counts), C.byref(self._nq)))
        return self._result()

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            load_library().cgvec_stream_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ServeSession:
    """cgvec_serve_*: batch-1 searches served by a RESIDENT scan kernel (no launch per query).  search() is the drop-in for
    Index.search(q, k) with one query; submit_device()/wait() pipeline device-resident queries."""

    def __init__(self, index: "Index", k: int, metric: int = COSINE, **options):
        self._ix, self.k, self.metric = index, k, metric
        self._h = C.c_void_p()
        _check(load_library().cgvec_serve_open(index._h, k, metric, C.byref(self._h)))
        for key, v in options.items():
            self.set(key, v)
        self._rows = np.empty(k, np.uint64); self._scores = np.empty(k, np.float32); self._count = np.zeros(1, np.uint32)
        self._ids = np.zeros((k, 16), np.uint8)
        self._p = (_ptr(self._rows), _ptr(self._ids), _ptr(self._scores), _ptr(self._count))
        self._search = load_library().cgvec_serve_search
        self._submit = load_library().cgvec_serve_submit
        self._t = C.c_uint32(); self._tref = C.byref(self._t)
        self.rows, self.scores, self.count = self._rows, self._scores, self._count

    def set(self, key: str, value: int):
        _check(load_library().cgvec_serve_set(self._h, key.encode(), int(value)))

    def search(self, query, want_ids: bool = False):
        """-> (rows u64[k], scores f32[k], count[, ids u8[k,16]]) — views of buffers reused by the next call"""
        q = np.ascontiguousarray(query, np.float32).reshape(-1)
        if q.shape[0] != self._ix.dim:
            raise CgvecError(ERR_BAD_DIM, f"query dimension {q.shape[0]} != index dimension {self._ix.dim}")
        rc = load_library().cgvec_serve_search(self._h, q.ctypes.data_as(C.c_void_p), self._p[0], self._p[1] if want_ids else None, self._p[2], self._p[3])
        if rc:
            _check(rc)
        if want_ids:
            return self._rows, self._scores, int(self._count[0]), self._ids
        return self._rows, self._scores, int(self._count[0])

    def search_raw(self, q: np.ndarray):
        """search() without argument marshalling: `q` must be a C-contiguous float32 array of `dim` elements; results land in
        self.rows / self.scores / self.count (the calling convention a server loop uses)."""
        rc = self._search(self._h, q.ctypes.data, self._p[0], None, self._p[2], self._p[3])
        if rc:
            _check(rc)

    def submit_device(self, d_query: int, d_rows: int, d_scores: int, d_counts: int) -> int:
        """Device-resident query (qstride floats) and result buffers (raw device pointers); returns the ticket."""
        rc = self._submit(self._h, d_query, 1, d_rows, d_scores, d_counts, self._tref)
        if rc:
            _check(rc)
        return self._t.value

    def wait(self, ticket: int):
        rc = load_library().cgvec_serve_wait(self._h, ticket)
        if rc:
            _check(rc)

    def pause(self):
        _check(load_library().cgvec_serve_pause(self._h))

    def timer_start(self):
        """CUDA-event bracket on the session's launch stream (kernel launch, queries and kernel exit lie inside)."""
        _check(load_library().cgvec_serve_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        _check(load_library().cgvec_serve_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(load_library().cgvec_serve_stats(self._h, C.byref(a), C.byref(b)))
        return {"launches": int(a.value), "served": int(b.value)}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            load_library().cgvec_serve_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Index:
    def __init__(self, dim: int, dtype: int = F32, device: int = 0, rank: int = 0, world: int = 1,
                 nccl_unique_id: Optional[bytes] = None, row_offset: int = 0, devices: Optional[Sequence[int]] = None):
        L = load_library()
        self._h = C.c_void_p()
        self.dim, self.dtype, self.device, self.rank, self.world, self.row_offset = dim, dtype, device, rank, world, row_offset
        if devices is not None and len(devices) > 1:          # single-process multi-GPU index
            dev = (C.c_int * len(devices))(*devices)
            _check(L.cgvec_create(dim, dtype, dev, len(devices), C.byref(self._h)))
        elif world == 1 and rank == 0 and row_offset == 0:
            dev = (C.c_int * 1)(device)
            _check(L.cgvec_create(dim, dtype, dev, 1, C.byref(self._h)))
        else:
            buf = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id else None
            _check(L.cgvec_create_rank(dim, dtype, device, rank, world, buf, row_offset, C.byref(self._h)))

    @classmethod
    def from_env(cls, dim: int, dtype: int = F32, enable_gpu: bool = False) -> "Index":
        """cgvec_create_from_env: `enable_gpu` is PerformanceConfig.enable_gpu (config_manager.rs:362-364); CODEGRAPH_ENABLE_GPU
        overrides it, CODEGRAPH_B200_DEVICES picks the GPUs.  Raises CgvecError(ERR_DISABLED) when the switch is off."""
        L = load_library()
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.dim, self.dtype, self.device, self.rank, self.world, self.row_offset = dim, dtype, 0, 0, 1, 0
        _check(L.cgvec_create_from_env(dim, dtype, 1 if enable_gpu else 0, C.byref(self._h)))
        return self

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            rc = load_library().cgvec_destroy(self._h)
            if rc:                                             # sessions still open: the index stays alive
                _check(rc)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(load_library().cgvec_len(self._h))

    # -- write side
    def reserve(self, n: int):
        _check(load_library().cgvec_reserve(self._h, n))

    def add(self, rows, ids=None):
        L = load_library()
        idb = _ids_to_bytes(ids)
        if self.dtype == F32:
            r = np.ascontiguousarray(rows, np.float32)
            fn = L.cgvec_add
        else:
            r = np.asarray(rows)
            r = np.ascontiguousarray(r if r.dtype == np.uint16 else r.astype(np.float16).view(np.uint16))
            fn = L.cgvec_add_f16
        if r.ndim != 2 or r.shape[1] != self.dim:
            raise CgvecError(ERR_BAD_DIM, f"rows have shape {r.shape}, index dimension is {self.dim}")
        if idb is not None and len(idb) != len(r):
            raise CgvecError(ERR_BAD_ARG, "ids / rows length mismatch")
        _check(fn(self._h, _ptr(idb), _ptr(r), r.shape[0]))

    def normalize_rows(self):
        _check(load_library().cgvec_normalize_rows(self._h))

    def fill_synthetic(self, n: int, seed: int, unit_norm: bool = True):
        _check(load_library().cgvec_fill_synthetic(self._h, n, seed, 1 if unit_norm else 0))

    # -- read side
    def search(self, queries, k: int, metric: int = COSINE, formula: int = FORMULA_SIMD, path: int = PATH_AUTO,
               want_ids: bool = False):
        """-> (rows u64[nq,k], scores f32[nq,k], counts u32[nq][, ids u8[nq,k,16]])"""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise CgvecError(ERR_BAD_DIM, f"query dimension {q.shape[1]} != index dimension {self.dim}")
        nq = q.shape[0]
        rows = np.full((nq, max(k, 1)), 2**64 - 1, np.uint64)
        scores = np.zeros((nq, max(k, 1)), np.float32)
        counts = np.zeros(nq, np.uint32)
        ids = np.zeros((nq, max(k, 1), 16), np.uint8) if want_ids else None
        o = SearchOpts(C.sizeof(SearchOpts), metric, formula, path, None, 0)
        _check(load_library().cgvec_search_ex(self._h, _ptr(q), nq, k, C.byref(o), _ptr(rows), _ptr(ids), _ptr(scores), _ptr(counts)))
        rows, scores = rows[:, :k], scores[:, :k]
        if want_ids:
            return rows, scores, counts, ids[:, :k]
        return rows, scores, counts

    def make_search_buffers(self, nq: int, k: int):
        """Preallocated host outputs + cached C pointers for search_into (the low-overhead calling convention a server
        loop uses: no per-call allocation or marshalling beyond the query pointer)."""
        rows = np.full((nq, k), 2**64 - 1, np.uint64); scores = np.zeros((nq, k), np.float32); counts = np.zeros(nq, np.uint32)
        opts = SearchOpts(C.sizeof(SearchOpts), COSINE, FORMULA_SIMD, PATH_AUTO, None, 0)
        return {"rows": rows, "scores": scores, "counts": counts, "opts": opts, "nq": nq, "k": k,
                "p_rows": _ptr(rows), "p_scores": _ptr(scores), "p_counts": _ptr(counts), "p_opts": C.byref(opts)}

    def search_into(self, queries: np.ndarray, bufs, metric: int = COSINE):
        """cgvec_search_ex with host buffers; `queries` must be C-contiguous float32 [nq, dim]; results land in bufs."""
        bufs["opts"].metric = metric
        rc = load_library().cgvec_search_ex(self._h, queries.ctypes.data_as(C.c_void_p), bufs["nq"], bufs["k"], bufs["p_opts"],
                                            bufs["p_rows"], None, bufs["p_scores"], bufs["p_counts"])
        if rc:
            _check(rc)

    def search_device(self, d_queries: int, nq: int, k: int, d_rows: int, d_scores: int, d_counts: int,
                      metric: int = COSINE, stream: int = 0, path: int = PATH_AUTO):
        """Device-resident I/O (raw device pointers); asynchronous on `stream`."""
        o = SearchOpts(C.sizeof(SearchOpts), metric, FORMULA_SIMD, path, C.c_void_p(stream or None), 1)
        _check(load_library().cgvec_search_ex(self._h, C.c_void_p(d_queries), nq, k, C.byref(o), C.c_void_p(d_rows), None,
                                              C.c_void_p(d_scores), C.c_void_p(d_counts)))

    def get_row(self, local_row: int) -> np.ndarray:
        out = np.empty(self.dim, np.float32)
        _check(load_library().cgvec_get_row(self._h, local_row, _ptr(out)))
        return out

    def get_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.float32)
        _check(load_library().cgvec_get_rows(self._h, first, n, _ptr(out)))
        return out

    def get(self, node_id) -> Optional[np.ndarray]:
        b = _ids_to_bytes([node_id])
        out = np.empty(self.dim, np.float32)
        rc = load_library().cgvec_get(self._h, _ptr(b), _ptr(out))
        if rc == ERR_NOT_FOUND:
            return None
        _check(rc)
        return out

    def row_of_id(self, node_id) -> Optional[int]:
        b = _ids_to_bytes([node_id]); r = C.c_uint64()
        rc = load_library().cgvec_row_of_id(self._h, _ptr(b), C.byref(r))
        if rc == ERR_NOT_FOUND:
            return None
        _check(rc)
        return int(r.value)

    def rescore(self, query, local_rows, metric: int = COSINE, formula: int = FORMULA_SEQ) -> np.ndarray:
        q = np.ascontiguousarray(query, np.float32)
        if q.shape != (self.dim,):
            raise CgvecError(ERR_BAD_DIM, f"query dimension {q.shape} != index dimension {self.dim}")
        lr = np.ascontiguousarray(local_rows, np.uint64)
        out = np.empty(len(lr), np.float32)
        _check(load_library().cgvec_rescore(self._h, _ptr(q), _ptr(lr), len(lr), metric, formula, _ptr(out)))
        return out

    def distances_first(self, query, limit: int) -> np.ndarray:
        q = np.ascontiguousarray(query, np.float32)
        out = np.empty(max(min(limit, len(self)), 1), np.float32)
        n = C.c_uint64()
        _check(load_library().cgvec_distances_first(self._h, _ptr(q), limit, _ptr(out), C.byref(n)))
        return out[: n.value].copy()

    def quantize_i8(self):
        """ModelOptimizer::quantize_batch, 8 bits (optimization.rs:212-224, 268-274)."""
        _check(load_library().cgvec_quantize_i8(self._h))

    def codes_i8(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), np.uint8)
        _check(load_library().cgvec_get_codes_i8(self._h, first, n, _ptr(out)))
        return out

    def search_optimized(self, query, limit: int):
        """OptimizationResult::search_optimized (optimization.rs:63-150) -> (rows, scores)."""
        q = np.ascontiguousarray(query, np.float32)
        k = max(int(limit), 1)
        rows = np.empty(k, np.uint64); scores = np.empty(k, np.float32); cnt = C.c_uint32()
        _check(load_library().cgvec_search_i8(self._h, _ptr(q), limit, _ptr(rows), _ptr(scores), C.byref(cnt)))
        return rows[: cnt.value].copy(), scores[: cnt.value].copy()

    def save_flat(self, path: str):
        """memory.rs:241-310 save_to_mmap format."""
        _check(load_library().cgvec_save_flat(self._h, path.encode()))

    def load_flat(self, path: str) -> int:
        """memory.rs:312-374 load_from_mmap format; appends the rows."""
        n = C.c_uint64()
        _check(load_library().cgvec_load_flat(self._h, path.encode(), C.byref(n)))
        return int(n.value)

    def stream(self, max_batch: int, k: int, metric: int = COSINE, path: int = PATH_AUTO) -> "QueryStream":
        """Double-buffered streaming of host query batches (cgvec_stream_*; BASELINE config 5)."""
        return QueryStream(self, max_batch, k, metric, path)

    def stats(self) -> Stats:
        s = Stats()
        _check(load_library().cgvec_get_stats(self._h, C.byref(s)))
        return s

    def trace(self, max_entries: int = 16384) -> np.ndarray:
        """-> int64[n, 3]: kind (1 scan, 2 merge, 3 exchange), start ns, end ns of every launch traced so far."""
        out = np.zeros((max_entries, 3), np.uint64); n = C.c_uint32()
        _check(load_library().cgvec_get_trace(self._h, _ptr(out), max_entries, C.byref(n)))
        return out[: n.value].astype(np.int64)

    def set_option(self, key: str, value: int):
        _check(load_library().cgvec_set_option(self._h, key.encode(), value))


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(load_library().cgvec_nccl_unique_id(buf))
    return buf.raw


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    b, e = C.c_uint64(), C.c_uint64()
    _check(load_library().cgvec_shard_range(n, world, rank, C.byref(b), C.byref(e)))
    return int(b.value), int(e.value)


def multi_locate(n_devices: int, global_row: int) -> Tuple[int, int]:
    s, l = C.c_uint32(), C.c_uint64()
    _check(load_library().cgvec_multi_locate(n_devices, global_row, C.byref(s), C.byref(l)))
    return int(s.value), int(l.value)


def multi_local_count(n_devices: int, shard: int, n_rows: int) -> int:
    return int(load_library().cgvec_multi_local_count(n_devices, shard, n_rows))


def merge_topk_host(rows, scores, counts, k: int, ascending: bool = False):
    r = np.ascontiguousarray(rows, np.uint64); s = np.ascontiguousarray(scores, np.float32)
    c = np.ascontiguousarray(counts, np.uint32)
    parts = r.shape[0]
    assert r.shape == (parts, k) and s.shape == (parts, k) and c.shape == (parts,)
    orow = np.empty(max(k, 1), np.uint64); osc = np.empty(max(k, 1), np.float32); cnt = C.c_uint32()
    _check(load_library().cgvec_merge_topk_host(_ptr(r), _ptr(s), _ptr(c), parts, k, 1 if ascending else 0, _ptr(orow), _ptr(osc), C.byref(cnt)))
    return orow[: cnt.value].copy(), osc[: cnt.value].copy()


def prefetch_k_basic(limit: int) -> int:
    return int(load_library().cgvec_prefetch_k_basic(limit))


def prefetch_k_filtered(limit: int) -> int:
    return int(load_library().cgvec_prefetch_k_filtered(limit))


def normalize_scores(scores) -> np.ndarray:
    s = np.ascontiguousarray(scores, np.float32).copy()
    load_library().cgvec_normalize_scores(_ptr(s), s.size)
    return s


# -------------------------------------------------------------------------------------------------
# mirrors of the reference interfaces
# -------------------------------------------------------------------------------------------------
@dataclass
class CodeNode:
    """The two fields of codegraph_core::CodeNode (node.rs:4-16) the vector path reads."""
    id: _uuid.UUID = field(default_factory=_uuid.uuid4)
    embedding: Optional[Sequence[float]] = None


class ParallelVectorOps:
    """simd_ops.rs:343-383"""

    @staticmethod
    def parallel_batch_similarity(query, embeddings, device: int = 0) -> np.ndarray:
        """simd_ops.rs:347-358 with similarity_fn = adaptive_cosine_similarity: one score per embedding, in order
        (every score computed on the device in the reference's operation order)."""
        emb = np.ascontiguousarray(embeddings, np.float32)
        if emb.ndim != 2 or emb.shape[0] == 0:
            return np.zeros(0, np.float32)
        ix = Index(emb.shape[1], F32, device)
        try:
            ix.add(emb)
            return ix.rescore(query, np.arange(emb.shape[0], dtype=np.uint64), COSINE, FORMULA_SIMD)
        finally:
            ix.close()

    @staticmethod
    def parallel_normalize_vectors(vectors, device: int = 0) -> np.ndarray:
        """simd_ops.rs:386-419 -> normalize_avx2 per vector (returns the normalised copy; zero vectors stay zero)."""
        v = np.ascontiguousarray(vectors, np.float32)
        if v.ndim != 2 or v.shape[0] == 0:
            return v.copy()
        ix = Index(v.shape[1], F32, device)
        try:
            ix.add(v)
            ix.normalize_rows()
            return ix.get_rows(0, v.shape[0])
        finally:
            ix.close()

    @staticmethod
    def parallel_top_k_search(query, embeddings, k: int, device: int = 0) -> List[Tuple[int, float]]:
        emb = np.ascontiguousarray(embeddings, np.float32)
        if emb.ndim != 2 or emb.shape[0] == 0:
            return []
        ix = Index(emb.shape[1], F32, device)
        try:
            ix.add(emb)
            rows, scores, counts = ix.search(query, k, COSINE)
            return [(int(rows[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]
        finally:
            ix.close()


class B200VectorStore:
    """trait VectorStore (codegraph-core/src/traits.rs:11-16) over one GPU-resident index."""

    def __init__(self, dimension: int, dtype: int = F32, device: int = 0, devices: Optional[Sequence[int]] = None):
        self.index = Index(dimension, dtype, device, devices=devices)

    def store_embeddings(self, nodes: Sequence[CodeNode]) -> None:
        withemb = [n for n in nodes if n.embedding is not None]       # nodes without an embedding are skipped
        if not withemb:                                                # (graph_vector.rs:472-476)
            return
        rows = np.asarray([n.embedding for n in withemb], np.float32)
        self.index.add(rows, [n.id for n in withemb])

    def search_similar(self, query_embedding, limit: int) -> List[_uuid.UUID]:
        q = np.asarray(query_embedding, np.float32)
        if q.size == 0 or limit == 0:                                  # surreal_store.rs:62-64
            return []
        _, _, counts, ids = self.index.search(q, limit, COSINE, want_ids=True)
        return [_uuid.UUID(bytes=ids[0, i].tobytes()) for i in range(int(counts[0]))]

    def get_embedding(self, node_id: _uuid.UUID) -> Optional[List[float]]:
        r = self.index.get(node_id)
        return None if r is None else r.tolist()


class B200Backend:
    """trait SurrealVectorBackend (surreal_store.rs:11-22): ("nodes:<uuid>", cosine DISTANCE) ascending."""

    def __init__(self, store: B200VectorStore):
        self.store = store
        self.last_column = None

    def upsert_nodes(self, nodes: Sequence[CodeNode]) -> None:
        self.store.store_embeddings(nodes)

    def vector_knn(self, column: str, query_embedding, limit: int, ef_search: int = 0) -> List[Tuple[str, float]]:
        self.last_column = column
        q = np.asarray(query_embedding, np.float32)
        if q.size == 0 or limit == 0:
            return []
        _, scores, counts, ids = self.store.index.search(q, limit, COSINE, want_ids=True)
        return [(f"nodes:{_uuid.UUID(bytes=ids[0, i].tobytes())}", float(np.float32(1.0) - scores[0, i]))
                for i in range(int(counts[0]))]

    def get_node_embedding(self, node_id: _uuid.UUID):
        return self.store.get_embedding(node_id)


@dataclass
class SearchResult:
    node_id: _uuid.UUID
    score: float


class SemanticSearch:
    """search.rs:14-144: over-fetch through the trait, exact re-score (search.rs:519-533 arithmetic, on the
    device via cgvec_rescore), stable sort descending, truncate, min-max normalise."""

    def __init__(self, vector_store: B200VectorStore):
        self.vector_store = vector_store

    def search_by_embedding(self, query_embedding, limit: int) -> List[SearchResult]:
        prefetch_k = prefetch_k_basic(limit)                                       # search.rs:113
        ids = self.vector_store.search_similar(query_embedding, prefetch_k)        # search.rs:114-117
        if not ids:
            return []
        ix = self.vector_store.index
        local = [ix.row_of_id(i) for i in ids]
        raw = ix.rescore(query_embedding, local, COSINE, FORMULA_SEQ)              # search.rs:120-129, :207-217
        order = sorted(range(len(ids)), key=lambda j: (-(raw[j]) if raw[j] == raw[j] else np.inf, j))   # :132-136 (stable)
        order = order[:limit]
        norm = normalize_scores(raw[order])                                        # search.rs:138
        return [SearchResult(ids[j], float(s)) for j, s in zip(order, norm)]


@dataclass
class ChunkRecord:
    """The fields of a `chunks` row the candidate stage reads (schema/codegraph.surql:331-345): its id, its parent node
    (None for orphan chunks) and its embedding."""
    id: _uuid.UUID
    parent_node: Optional[_uuid.UUID]
    embedding: Sequence[float]


@dataclass
class ChunkCandidate:
    node_id: _uuid.UUID          # parent node of the chunk hit (<string> parent_node AS node_id, codegraph.surql:403)
    vector_score: float          # 1f - distance (codegraph.surql:412)
    chunk_id: _uuid.UUID


class ChunkCandidateStage:
    """Step 1-2 of fn::semantic_search_nodes_via_chunks (schema/codegraph.surql:318-417), the vector stage behind
    execute_semantic_code_search (codegraph-mcp-tools/src/graph_tool_executor.rs:578-591): the 100 nearest chunk
    embeddings (`<|100,200|>`), minus chunks without a parent node, cut to 3 x safe_limit rows (safe_limit = limit when
    1 <= limit <= 100, else 10), each mapped to (parent node, 1 - cosine distance).  The KNN itself is the EXACT
    brute-force scan of this library (B200Backend.vector_knn) in place of SurrealDB's approximate HNSW; BM25, the 0.9/0.1
    blend and the graph enrichment stay in SurrealDB."""

    KNN = 100                                   # the literal of `<|100,200|>`

    def __init__(self, dimension: int, dtype: int = F32, device: int = 0, devices: Optional[Sequence[int]] = None):
        self.store = B200VectorStore(dimension, dtype, device, devices)
        self.backend = B200Backend(self.store)
        self.parent: dict = {}                  # chunk id -> parent node id | None

    def upsert_chunks(self, chunks: Sequence[ChunkRecord]) -> None:
        self.backend.upsert_nodes([CodeNode(c.id, c.embedding) for c in chunks])
        for c in chunks:
            self.parent[c.id] = c.parent_node

    def candidates(self, query_embedding, limit: int) -> List[ChunkCandidate]:
        safe_limit = limit if 0 < limit <= 100 else 10                              # codegraph.surql:325
        chunk_limit = safe_limit * 3                                                # :326
        hits = self.backend.vector_knn(f"embedding_{self.store.index.dim}", query_embedding, self.KNN)   # :334-339, ascending distance
        out: List[ChunkCandidate] = []
        for rid, distance in hits:
            cid = _uuid.UUID(rid.split(":", 1)[1])
            parent = self.parent.get(cid)
            if parent is None:                                                      # parent_node != NONE
                continue
            out.append(ChunkCandidate(parent, float(np.float32(1.0) - np.float32(distance)), cid))
            if len(out) == chunk_limit:                                             # LIMIT $chunk_limit
                break
        return out


def resolve_symbol(index: "Index", target_embedding, candidate_rows, threshold: float = 0.75):
    """AI symbol resolver arg-max (codegraph-mcp/src/indexer.rs:2827-2843, cosine :2965-2979): among the candidate rows
    that survived the caller's trigram prefilter, the FIRST row with the highest cosine (search.rs:519-533 arithmetic,
    computed on the device) strictly above `threshold`; None if none.  -> (row, similarity) | None"""
    cand = np.ascontiguousarray(candidate_rows, np.uint64)
    if cand.size == 0:
        return None
    sims = index.rescore(target_embedding, cand, COSINE, FORMULA_SEQ)
    best = None
    for r, s in zip(cand.tolist(), sims.tolist()):
        if s > threshold and (best is None or s > best[1]):      # indexer.rs:2832-2840: strict >, first best wins
            best = (int(r), float(s))
    return best


def resolve_symbols(index: "Index", target_embeddings, threshold: float = 0.75):
    """The resolver's arg-max for U unresolved references at once, over ALL symbol rows of `index` instead of a trigram
    prefilter (codegraph-mcp/src/indexer.rs:2673-2878 runs the U loop with rayon, :1914,1983).  With U >= 8 (16 for f32
    storage) and >= 32768 symbols this is ONE dense U x S pass on the tensor cores per 256 references (k = 1, survivors
    re-scored in the sequential cosine of search.rs:519-533 / indexer.rs:2965-2979, exactness proven on the device);
    smaller problems take one exact scan per reference.  Ties go to the lower row = the reference's "first best wins"
    in row order; entries at or below `threshold` become None.  -> [ (row, similarity) | None ] * U"""
    t = np.ascontiguousarray(target_embeddings, np.float32)
    if t.ndim == 1:
        t = t[None, :]
    if t.shape[0] == 0 or len(index) == 0:
        return [None] * t.shape[0]
    rows, sims, counts = index.search(t, 1, COSINE, formula=FORMULA_SEQ)
    return [(int(rows[u, 0]), float(sims[u, 0])) if counts[u] and sims[u, 0] > threshold else None for u in range(t.shape[0])]


class GpuAcceleration:
    """gpu.rs:109-381 API shape: upload_vectors(flat, dimension) -> data handle; compute_distances(query, data, limit)."""

    def __init__(self, device: int = 0):
        self.device = device

    def upload_vectors(self, vectors, dimension: int) -> Index:
        flat = np.ascontiguousarray(vectors, np.float32).reshape(-1)
        if flat.size % dimension != 0:                                             # gpu.rs:226-230
            raise CgvecError(ERR_BAD_ARG, "Vector data length not divisible by dimension")
        ix = Index(dimension, F32, self.device)
        ix.add(flat.reshape(-1, dimension))
        return ix

    def compute_distances(self, query, gpu_data: Index, limit: int) -> np.ndarray:
        q = np.asarray(query, np.float32)
        if q.size != gpu_data.dim:                                                 # gpu.rs:265-271
            raise CgvecError(ERR_BAD_DIM, f"Query dimension {q.size} doesn't match GPU data dimension {gpu_data.dim}")
        return gpu_data.distances_first(q, limit)
