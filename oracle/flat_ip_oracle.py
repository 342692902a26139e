"""CPU oracle for the flat inner-product top-k search.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (ance_b200/) never does.

PARITY UNPINNED UPSTREAM: the arithmetic restated here lives in an un-vendored, *unpinned*
dependency (faiss-cpu, reference setup.py:22) that is not installed in this image and cannot be
(no network); the reference has no tests or golden vectors for it.  What is
restated is faiss's published IndexFlatIP contract, anchored on the reference's call sites:
  drivers/run_ann_data_gen.py:269-271   cpu_index = faiss.IndexFlatIP(dim); cpu_index.add(passage_embedding)
  drivers/run_ann_data_gen.py:276,303   _, I = cpu_index.search(query_embedding, k)
  drivers/run_ann_data_gen_dpr.py:238-252 (same, k = 100 / topk_training)
Contract: for each query the k rows of the index with the largest fp32 inner product, sorted by
score descending; labels are int64 row numbers in insertion order; if the index holds fewer than k
rows the tail is label -1 with score = lowest float.  faiss computes scores with a blocked BLAS
sgemm (fp32 accumulate, summation order unspecified) and leaves the order of equal scores
unspecified.  This oracle fixes both degrees of freedom:
  * canonical score = the fp32 inputs' dot product accumulated in fp64, rounded once to fp32
    (within 1 ulp of any fp32 summation order's exact value; what libance_b200 emits bit for bit);
  * ties are ordered by the smaller row number.
`near_ties()` reports the queries whose k-th / (k+1)-th canonical scores are closer than fp32
summation noise, i.e. where a real faiss run could legitimately return a different set.
"""
from __future__ import annotations

import numpy as np

LOWEST = np.finfo(np.float32).min


def search_bruteforce(P: np.ndarray, Q: np.ndarray, k: int):
    """Definition-level oracle (small inputs): full fp64 score matrix, stable sort."""
    P = np.ascontiguousarray(P, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    nq, n = Q.shape[0], P.shape[0]
    D = np.full((nq, k), LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if n == 0 or nq == 0:
        return D, I
    S = (Q.astype(np.float64) @ P.astype(np.float64).T).astype(np.float32)
    order = np.argsort(-S, axis=1, kind="stable")[:, :k]  # stable: equal scores keep ascending row order
    kk = order.shape[1]
    I[:, :kk] = order
    D[:, :kk] = np.take_along_axis(S, order, axis=1)
    return D, I


def search(P: np.ndarray, Q: np.ndarray, k: int, slack: int = 64, q_block: int = 512, p_block: int = 262144):
    """Blocked oracle for sizes where the full fp64 matrix is too large.

    Pass 1 is the faiss-like arithmetic (blocked fp32 sgemm through BLAS) keeping k + slack
    candidates per query; pass 2 recomputes the candidates canonically (fp64 accumulate) and
    orders them by (score desc, row asc).  The slack is verified: the weakest kept fp32 score must
    lie below the k-th canonical score by more than fp32 summation noise.
    """
    P = np.ascontiguousarray(P, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    nq, n = Q.shape[0], P.shape[0]
    D = np.full((nq, k), LOWEST, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    if n == 0 or nq == 0:
        return D, I
    keep = min(n, k + slack)
    pn = float(np.sqrt((P.astype(np.float64) ** 2).sum(1)).max())
    for q0 in range(0, nq, q_block):
        q = Q[q0:q0 + q_block]
        cs = np.full((q.shape[0], 0), 0, dtype=np.float32)
        ci = np.full((q.shape[0], 0), 0, dtype=np.int64)
        for p0 in range(0, n, p_block):
            s = q @ P[p0:p0 + p_block].T  # fp32 BLAS sgemm, as faiss
            kk = min(keep, s.shape[1])
            part = np.argpartition(-s, kk - 1, axis=1)[:, :kk]
            cs = np.concatenate([cs, np.take_along_axis(s, part, axis=1)], axis=1)
            ci = np.concatenate([ci, part.astype(np.int64) + p0], axis=1)
            if cs.shape[1] > keep:
                sel = np.argpartition(-cs, keep - 1, axis=1)[:, :keep]
                cs = np.take_along_axis(cs, sel, axis=1)
                ci = np.take_along_axis(ci, sel, axis=1)
        for r in range(q.shape[0]):
            rows = ci[r]
            exact = (P[rows].astype(np.float64) @ q[r].astype(np.float64)).astype(np.float32)
            order = np.lexsort((rows, -exact.astype(np.float64)))
            kk = min(k, rows.shape[0])
            top = order[:kk]
            I[q0 + r, :kk] = rows[top]
            D[q0 + r, :kk] = exact[top]
            if keep < n and kk == k:
                # fp32 sgemm noise bound for a d-term dot product: d * 2^-24 * |q| * |p|
                noise = P.shape[1] * 2.0 ** -24 * float(np.linalg.norm(q[r].astype(np.float64))) * pn
                weakest = float(cs[r].min())
                if not weakest + 2 * noise < float(exact[top[-1]]):
                    raise AssertionError(
                        f"oracle slack too small for query {q0 + r}: weakest kept fp32 score {weakest} vs "
                        f"k-th canonical {float(exact[top[-1]])}; raise slack")
    return D, I


def canonical_scores(P: np.ndarray, Q: np.ndarray, I: np.ndarray) -> np.ndarray:
    """Canonical (fp64-accumulated, fp32-rounded) scores of the labelled rows; -1 labels -> lowest."""
    out = np.full(I.shape, LOWEST, dtype=np.float32)
    for r in range(I.shape[0]):
        ok = I[r] >= 0
        out[r, ok] = (P[I[r, ok]].astype(np.float64) @ Q[r].astype(np.float64)).astype(np.float32)
    return out


def near_ties(P: np.ndarray, Q: np.ndarray, k: int, D: np.ndarray, extra: int = 8):
    """Queries whose k-th and (k+1)-th canonical scores differ by less than fp32 summation noise.
    Returns a dict {query: gap}.  Uses one extra blocked search with k + extra."""
    D2, _ = search(P, Q, k + extra)
    pn = float(np.sqrt((P.astype(np.float64) ** 2).sum(1)).max())
    out = {}
    for r in range(Q.shape[0]):
        gap = float(D2[r, k - 1]) - float(D2[r, k]) if D2.shape[1] > k else np.inf
        noise = P.shape[1] * 2.0 ** -24 * float(np.linalg.norm(Q[r].astype(np.float64))) * pn
        if gap < 2 * noise:
            out[r] = gap
    return out


def merge_shards(Ds, Is, k: int):
    """Reference precedent for the sharded form: utils/eval_mrr.py:175-183 (concatenate per-rank
    top-k, re-sort).  Ordering (score desc, label asc); -1 labels are padding."""
    D = np.concatenate(Ds, axis=1)
    I = np.concatenate(Is, axis=1)
    nq = D.shape[0]
    Do = np.full((nq, k), LOWEST, dtype=np.float32)
    Io = np.full((nq, k), -1, dtype=np.int64)
    for r in range(nq):
        ok = I[r] >= 0
        d, i = D[r, ok], I[r, ok]
        order = np.lexsort((i, -d.astype(np.float64)))[:k]
        Do[r, :order.shape[0]] = d[order]
        Io[r, :order.shape[0]] = i[order]
    return Do, Io


def search_c(P: np.ndarray, Q: np.ndarray, k: int):
    """The C restatement (oracle/flat_ip_oracle.c, built by ance_b200.build.build_oracle): plain loops, fp64
    accumulation, heap selection.  Independent of BLAS and of numpy's sort; agrees with search_bruteforce()."""
    import ctypes as C
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle.so")
    if not os.path.exists(so):
        from ance_b200.build import build_oracle   # compiling the checker is not using it in the product
        build_oracle()
    lib = C.CDLL(so)
    lib.flat_ip_search_exact.restype = C.c_int
    lib.flat_ip_search_exact.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    P = np.ascontiguousarray(P, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    nq, dim = Q.shape[0], (Q.shape[1] if Q.ndim == 2 else P.shape[1])
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    if lib.flat_ip_search_exact(P.ctypes.data, P.shape[0], Q.ctypes.data, nq, dim, k, D.ctypes.data, I.ctypes.data) != 0:
        raise MemoryError("flat_ip_search_exact: allocation failed")
    return D, I
