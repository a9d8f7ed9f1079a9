#!/usr/bin/env python
"""Bring-up aid for the tcgen05 path: small cases first, prints stats and the first mismatch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import oracle
cg = ge.load_package()

def case(n, d, nq, k, seed=0, **opts):
    rng = np.random.default_rng(seed)
    rows = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    qs = rng.standard_normal((nq, d)).astype(np.float32)
    ref = rows.astype(np.float16).astype(np.float32)
    ix = cg.Index(d, cg.F16)
    ix.add(rows)
    for kk, v in opts.items():
        ix.set_option(kk, v)
    t = time.time()
    try:
        r, s, c = ix.search(qs, k, cg.COSINE, path=cg.PATH_TENSOR)
    except cg.CgvecError as e:
        print(f"n={n} d={d} nq={nq} k={k}: ERROR {e}"); ix.close(); return False
    dt = time.time() - t
    st = ix.stats()
    bad = 0
    for qi in range(nq):
        wi, ws = oracle.parallel_top_k_search(qs[qi], ref, k)
        if r[qi, :len(wi)].tolist() != wi.tolist() or not np.array_equal(s[qi, :len(wi)], ws):
            if bad == 0:
                print("  first mismatch q", qi, "got", r[qi, :6], s[qi, :4], "want", wi[:6], ws[:4])
            bad += 1
    print(f"n={n} d={d} nq={nq} k={k}: batches={st.tc_batches} fallbacks={st.tc_fallbacks} mismatched_queries={bad} launches={st.kernel_launches} {dt*1e3:.1f} ms", flush=True)
    ix.close()
    return bad == 0

if __name__ == "__main__":
    case(2000, 64, 16, 5)
    case(20000, 64, 16, 5)
    case(20000, 128, 16, 10)
    case(100000, 128, 64, 10)
    case(100000, 768, 64, 10)
    case(200000, 1024, 64, 100)
