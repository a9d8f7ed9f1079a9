#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.json.

The reference (Rust) cannot be executed in the build image, so these fixtures are produced by the CPU oracle
(oracle/cgvec_oracle.c) from the reference's OWN test recipes — the inputs are exactly what the reference tests build
(codegraph-vector/src/simd_ops.rs:461-472, codegraph-vector/tests/model_optimization_tests.rs:36-58,383-424,
codegraph-vector/src/search.rs:178-205) — and are committed so that (a) drift of the oracle itself is caught on CPU and
(b) the GPU path is checked against fixed bytes, not only against a freshly computed oracle.  Floats are stored as
hex bit patterns.  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402


def hexf(a):
    return [format(int(x), "08x") for x in np.asarray(a, np.float32).view(np.uint32)]


def main():
    out = {}
    # (1) simd_ops.rs:461-472 test_parallel_operations
    q = np.ones(256, np.float32)
    rows = (np.arange(1000)[:, None] + np.arange(256)[None, :]).astype(np.float32)
    i, s = oracle.parallel_top_k_search(q, rows, 10)
    out["parallel_operations"] = {"n": 1000, "d": 256, "k": 10, "indices": i.tolist(), "scores": hexf(s)}
    # (2) model_optimization_tests.rs:36-58 vectors, query = row 0
    vecs = oracle.generate_optimization_vectors(1000, 128, 11223)
    out["optimization_vectors"] = {"count": 1000, "dim": 128, "seed": 11223,
                                   "sha256_f32": hashlib.sha256(vecs.tobytes()).hexdigest(), "row0_head": hexf(vecs[0, :8])}
    q = vecs[0]
    i, s = oracle.parallel_top_k_search(q, vecs, 10)
    out["optimization_simd_top10"] = {"indices": i.tolist(), "scores": hexf(s)}
    i, d = oracle.search_baseline(q, vecs, 10)
    out["optimization_search_baseline_top10"] = {"indices": i.tolist(), "distances": hexf(d)}
    i, s = oracle.inmemory_search_similar(q, vecs, 10)
    out["optimization_inmemory_top10"] = {"indices": i.tolist(), "scores": hexf(s)}
    codes = oracle.quantize_batch_u8(vecs)
    i, s = oracle.search_optimized_i8(q, codes, 10)
    out["optimization_int8"] = {"codes_sha256": hashlib.sha256(codes.tobytes()).hexdigest(), "indices": i.tolist(), "scores": hexf(s)}
    # (3) search.rs:178-205 deterministic text embedding
    out["hash_text_embedding"] = {t: {"head": hexf(oracle.hash_text_embedding(t, 384)[:8]),
                                      "sha256": hashlib.sha256(oracle.hash_text_embedding(t, 384).tobytes()).hexdigest()}
                                  for t in ("fn main() {}", "__query__", "")}
    # (4) synthetic bench inputs
    sr = oracle.synth_rows(0xC0DE6A9F, 0, 256, 768, True, False)
    sq = oracle.synth_rows(0x5EED0001, 0, 4, 768, True, False)
    out["synthetic"] = {"rows_seed": "0xC0DE6A9F", "queries_seed": "0x5EED0001", "rows_sha256_first256x768": hashlib.sha256(sr.tobytes()).hexdigest(),
                        "queries_sha256_first4x768": hashlib.sha256(sq.tobytes()).hexdigest()}
    big = oracle.synth_rows(0xC0DE6A9F, 0, 10_000, 768, True, False)      # BASELINE config 1: 10k x 768, 1 query, top-10
    i, s = oracle.parallel_top_k_search(sq[0], big, 10)
    out["config1_10k_x_768"] = {"indices": i.tolist(), "scores": hexf(s)}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
