"""Host-side post-processing of one ANN refresh: negatives, dev NDCG, output files.

Same results as the reference's pure-Python loops
  * GenerateNegativePassaageID   drivers/run_ann_data_gen.py:339-396
  * EvalDevQuery                 drivers/run_ann_data_gen.py:399-440
  * the writers                  drivers/run_ann_data_gen.py:315-334 (data file first, ann_ndcg json last:
                                 the trainer discovers a version by the json, utils/util.py:229-243)
but vectorised with numpy (the O(nq*k) Python loop with `in list` membership is the longest serial
segment of a refresh once encode and search run on the GPU).

Sampling order.  The reference visits each query's k neighbours in an order drawn from the
module-level, unseeded `random`.  `sampler="reference"` consumes Python's `random` exactly as the
reference does (same calls, same order), so under `random.seed(s)` the output files are byte-identical
to the reference's; `sampler="fast"` draws all permutations at once from a numpy Generator.
"""
from __future__ import annotations

import json
import math
import os
import random
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np


_FAST_CHUNK = 16384   # queries per block of the "fast" sampler


def _first_occurrence(pids: np.ndarray) -> np.ndarray:
    """mask[r, j] = True iff pids[r, j] does not appear in pids[r, :j]."""
    n, m = pids.shape
    order = np.argsort(pids, axis=1, kind="stable")
    sp = np.take_along_axis(pids, order, axis=1)
    first_sorted = np.ones((n, m), dtype=bool)
    first_sorted[:, 1:] = sp[:, 1:] != sp[:, :-1]
    first = np.empty((n, m), dtype=bool)
    np.put_along_axis(first, order, first_sorted, axis=1)
    return first


def generate_negatives(query_embedding2id: np.ndarray, passage_embedding2id: np.ndarray, positives: Dict[int, int],
                       I: np.ndarray, negative_sample: int, select_topk: bool = False, sampler: str = "fast",
                       seed: Optional[int] = None, as_arrays: bool = False, threads: int = 8):
    """-> ({qid: [neg pids]}, mrr_sum, num_queries).  `I` holds global passage ROWS (faiss labels);
    every query of `query_embedding2id` is effective (the reference builds effective_q_id from the same
    array, run_ann_data_gen.py:305).
    as_arrays: -> ((neg [nq, negative_sample] int64 padded with -1, counts [nq]), mrr_sum, num_queries) instead of the
    dict — what `write_training_data_arrays` consumes without creating nq Python lists."""
    nq, k = I.shape
    qids = np.asarray(query_embedding2id).reshape(-1).astype(np.int64)
    if nq == 0:
        return ((np.empty((0, negative_sample), np.int64), np.empty(0, np.int64)) if as_arrays else {}), 0.0, 0
    pos = np.fromiter((positives[int(q)] for q in qids), dtype=np.int64, count=nq)  # KeyError like the reference
    p2id = np.asarray(passage_embedding2id).reshape(-1)
    if select_topk:
        sel = I[:, :negative_sample + 1]
    elif sampler == "reference":
        perm = np.empty((nq, k), dtype=np.int64)
        base = list(range(k))
        for r in range(nq):
            p = base.copy()
            random.shuffle(p)
            perm[r] = p
        sel = np.take_along_axis(I, perm, axis=1)
    elif sampler == "fast":
        # Only the head of the shuffled list is ever read (the loop stops after negative_sample valid candidates), so
        # draw the first m positions of a uniform random permutation -- the m smallest of k i.i.d. keys, in key order --
        # instead of shuffling all k; rows that run out of valid candidates within m fall back to a full permutation.
        # Row blocks are independent (their own generator, seeded by (seed, block)): they run on a few threads, numpy
        # releases the GIL inside random / argpartition / sort / take.
        m = min(k, max(negative_sample + 16, 32))
        mat = np.full((nq, negative_sample), -1, dtype=np.int64)
        counts = np.zeros(nq, dtype=np.int64)
        blocks = list(range(0, nq, _FAST_CHUNK))
        entropy = np.random.SeedSequence().entropy if seed is None else seed

        def work(bi: int) -> float:
            lo = blocks[bi]
            hi = min(nq, lo + _FAST_CHUNK)
            rng = np.random.default_rng([entropy, bi]) if len(blocks) > 1 else np.random.default_rng(seed if seed is not None else entropy)
            keys = rng.random((hi - lo, k), dtype=np.float32)
            Ic, pc = I[lo:hi], pos[lo:hi]
            if m < k:
                part = np.argpartition(keys, m - 1, axis=1)[:, :m]
                head = np.take_along_axis(part, np.argsort(np.take_along_axis(keys, part, axis=1), axis=1), axis=1)
            else:
                head = np.argsort(keys, axis=1)
            sel = np.take_along_axis(Ic, head, axis=1)
            short = np.zeros(hi - lo, dtype=bool)
            if m < k:
                pids_h = p2id[np.where(sel < 0, 0, sel)]
                # (<=: the reference's break needs one more valid candidate than it keeps)
                short = ((_first_occurrence(pids_h) & (pids_h != pc[:, None])).sum(axis=1) <= negative_sample) | (sel < 0).any(axis=1)
            mrr = 0.0
            if short.any():
                rows = np.nonzero(short)[0]
                full = np.argsort(keys[rows], axis=1)
                m2, c2, r2 = _negatives_arrays(pc[rows], p2id, np.take_along_axis(Ic[rows], full, axis=1), negative_sample)
                mat[lo + rows], counts[lo + rows] = m2, c2
                keep = np.nonzero(~short)[0]
                m1, c1, r1 = _negatives_arrays(pc[keep], p2id, sel[keep], negative_sample)
                mat[lo + keep], counts[lo + keep] = m1, c1
                mrr = r1 + r2
            else:
                mat[lo:hi], counts[lo:hi], mrr = _negatives_arrays(pc, p2id, sel, negative_sample)
            return mrr

        if len(blocks) > 1 and threads > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(threads, len(blocks))) as ex:
                mrr = float(sum(ex.map(work, range(len(blocks)))))
        else:
            mrr = float(sum(work(bi) for bi in range(len(blocks))))
        if as_arrays:
            return (mat, counts), mrr, nq
        return _dict_from_arrays(qids, mat, counts), mrr, nq
    else:
        raise ValueError(f"unknown sampler {sampler!r}")
    mat, counts, mrr = _negatives_arrays(pos, p2id, sel, negative_sample)
    if as_arrays:
        return (mat, counts), mrr, nq
    return _dict_from_arrays(qids, mat, counts), mrr, nq


def _dict_from_arrays(qids: np.ndarray, mat: np.ndarray, counts: np.ndarray) -> Dict[int, List[int]]:
    if mat.shape[0] and (counts == mat.shape[1]).all():   # the common case: one conversion instead of nq slices
        return dict(zip(qids.tolist(), mat.tolist()))
    return {int(q): mat[r, :counts[r]].tolist() for r, q in enumerate(qids.tolist())}


def _negatives_arrays(pos: np.ndarray, p2id: np.ndarray, sel: np.ndarray, negative_sample: int):
    """The reference's scan (run_ann_data_gen.py:357-385) over the candidate rows `sel` [nq, m] in the given order.
    -> (neg pids [nq, negative_sample] padded with -1, counts [nq], mrr_sum)."""
    nq = sel.shape[0]
    mat = np.full((nq, negative_sample), -1, dtype=np.int64)
    if nq == 0:
        return mat, np.zeros(0, dtype=np.int64), 0.0
    if (sel < 0).any():
        raise IndexError("search returned -1 labels (fewer rows than k); the reference would index "
                         "passage_embedding2id[-1] silently — refusing")
    pids = p2id[sel]
    is_pos = pids == pos[:, None]
    valid = _first_occurrence(pids) & ~is_pos
    csum = np.cumsum(valid, axis=1)
    take = valid & (csum <= negative_sample)
    # the reference's loop breaks at the first valid candidate met once negative_sample were taken
    brk_mask = valid & (csum == negative_sample + 1)
    m = sel.shape[1]
    brk = np.where(brk_mask.any(axis=1), brk_mask.argmax(axis=1), m)
    ranks = np.arange(1, m + 1)[None, :]
    hit = is_pos & (ranks <= 10) & (np.arange(m)[None, :] < brk[:, None])
    mrr = float((hit / ranks).sum())
    counts = take.sum(axis=1)
    r_idx, c_idx = np.nonzero(take)
    mat[r_idx, csum[r_idx, c_idx] - 1] = pids[r_idx, c_idx]
    return mat, counts.astype(np.int64), mrr


def _negatives_from_selection(qids: np.ndarray, pos: np.ndarray, p2id: np.ndarray, sel: np.ndarray,
                              negative_sample: int) -> Tuple[Dict[int, List[int]], float, int]:
    mat, counts, mrr = _negatives_arrays(pos, p2id, sel, negative_sample)
    return _dict_from_arrays(np.asarray(qids), mat, counts), mrr, sel.shape[0]


def ndcg_cut(ranked_pids: Sequence[int], qrel: Dict[int, int], cut: int = 10) -> float:
    """trec_eval's ndcg_cut_k (what the reference reads from pytrec_eval, run_ann_data_gen.py:426-435):
    gain = relevance, discount = log2(rank + 1), ideal from all judged documents of the query."""
    dcg = 0.0
    for r, pid in enumerate(ranked_pids[:cut], start=1):
        g = qrel.get(int(pid), 0)
        if g > 0:
            dcg += g / math.log2(r + 1)
    ideal = sorted((g for g in qrel.values() if g > 0), reverse=True)[:cut]
    idcg = sum(g / math.log2(r + 1) for r, g in enumerate(ideal, start=1))
    return dcg / idcg if idcg > 0 else 0.0


def eval_dev_query(query_embedding2id: np.ndarray, passage_embedding2id: np.ndarray,
                   dev_query_positive_id: Dict[int, Dict[int, int]], I: np.ndarray) -> Tuple[float, int]:
    """Mean ndcg_cut_10 over the dev queries that have qrels, from the first 50 neighbours with
    pid de-duplication (run_ann_data_gen.py:413-423).  -> (ndcg, evaluated query count)."""
    top = I[:, :50]
    if top.size and (top < 0).any():
        raise IndexError("search returned -1 labels (fewer rows than k)")
    pids = np.asarray(passage_embedding2id).reshape(-1)[top]
    first = _first_occurrence(pids) if pids.size else np.zeros_like(pids, dtype=bool)
    pred: Dict[int, np.ndarray] = {}
    qids = np.asarray(query_embedding2id).reshape(-1)
    for r in range(top.shape[0]):
        pred[int(qids[r])] = pids[r, first[r]]  # duplicate qids: last wins, as the dict assignment does
    total, n = 0.0, 0
    for qid, ranked in pred.items():
        qrel = dev_query_positive_id.get(qid)
        if qrel is None:
            continue
        n += 1
        total += ndcg_cut(ranked, qrel, 10)
    if n == 0:
        raise ZeroDivisionError("no dev query has a qrel (the reference divides by zero here too)")
    return total / n, n


def write_training_data(path: str, query_embedding2id: np.ndarray, positives: Dict[int, int],
                        negatives: Dict[int, List[int]], sampler: str = "fast", seed: Optional[int] = None) -> int:
    """`qid \\t pos_pid \\t n1,n2,...` per query, in shuffled query order (run_ann_data_gen.py:318-329).
    Returns the number of lines."""
    qids = np.asarray(query_embedding2id).reshape(-1)
    order = list(range(len(qids)))
    if sampler == "reference":
        random.shuffle(order)
    else:
        order = np.random.default_rng(None if seed is None else seed + 1).permutation(len(qids)).tolist()
    n = 0
    tmp = staging_path(path)
    with open(tmp, "w") as f:
        for qi in order:
            qid = int(qids[qi])
            if qid not in positives or qid not in negatives:
                continue
            f.write("{}\t{}\t{}\n".format(qid, positives[qid], ",".join(str(p) for p in negatives[qid])))
            n += 1
    os.replace(tmp, path)  # a reader never sees a partial file (the reference writes in place)
    return n


def write_training_data_arrays(path: str, query_embedding2id: np.ndarray, positives: Dict[int, int], neg: np.ndarray,
                               counts: np.ndarray, seed: Optional[int] = None) -> int:
    """write_training_data for the array form of the negatives (generate_negatives(as_arrays=True)): the same lines in
    the same ("fast" sampler) order, formatted and written by libance_b200's host code instead of nq Python joins."""
    import ctypes as C
    from . import _lib
    qids = np.ascontiguousarray(np.asarray(query_embedding2id).reshape(-1), dtype=np.int64)
    n = len(qids)
    pos = np.fromiter((positives[int(q)] for q in qids), dtype=np.int64, count=n)
    order = np.ascontiguousarray(np.random.default_rng(None if seed is None else seed + 1).permutation(n), dtype=np.int64)
    neg = np.ascontiguousarray(neg, dtype=np.int64)
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    tmp = staging_path(path)
    written = C.c_int64()
    _lib.check(_lib.load().ance_write_training_data_host(tmp.encode(), qids.ctypes.data, pos.ctypes.data, neg.ctypes.data,
                                                         counts.ctypes.data, order.ctypes.data, n, neg.shape[1],
                                                         C.byref(written)))
    os.replace(tmp, path)
    return int(written.value)


def write_ndcg(path: str, ndcg: float, checkpoint: str) -> None:
    """run_ann_data_gen.py:331-334 — written AFTER the data file."""
    write_json_atomic(path, {"ndcg": ndcg, "checkpoint": checkpoint})


def staging_path(path: str) -> str:
    """Temp name for an atomic write into the trainer-polled output_dir.  The unmodified trainer parses EVERY file
    whose name starts with `ann_ndcg_` as `int(name[len('ann_ndcg_'):])` (utils/util.py:227-236, polled from
    run_ann.py:184), so the staged file must not carry that prefix: `.tmp.<name>.<pid>` in the same directory
    (same filesystem, so os.replace stays atomic)."""
    d, base = os.path.split(path)
    return os.path.join(d, ".tmp.{}.{}".format(base, os.getpid()))


def write_json_atomic(path: str, obj) -> None:
    tmp = staging_path(path)
    with open(tmp, "w") as f:
        json.dump(obj, f)
    os.replace(tmp, path)


def query_chunk(num_queries: int, output_num: int, chunk_factor: int) -> Tuple[int, int]:
    """run_ann_data_gen.py:281-296.  The reference evaluates `output_num % chunk_factor` before its
    `chunk_factor <= 0` guard (ZeroDivisionError for 0); here 0 / negative mean 'one chunk'."""
    if chunk_factor <= 0:
        chunk_factor = 1
    effective = output_num % chunk_factor
    per = num_queries // chunk_factor
    start = per * effective
    end = num_queries if effective == chunk_factor - 1 else start + per
    return start, end
