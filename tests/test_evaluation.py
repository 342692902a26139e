"""Offline metrics (ance_b200/evaluation.py = cells 8-13 of the reference's `evaluation/Calculate Metrics.ipynb`).

* MS MARCO MRR@10 is pinned by tests/golden/msmarco_mrr.json, produced by the reference's own utils/msmarco_eval.py.
* EvalDevQuery's bookkeeping (pid de-duplication, hole rates, the 1000-slot candidate lists) is checked against a literal
  restatement of the notebook loop; the trec_eval measures it reads from pytrec_eval (absent here: parity unpinned) are
  restated from trec_eval's definitions by brute force.
* rerank / load_dumps: properties on small random data (CPU); the GPU full-rank path is in tests/test_gpu_driver.py."""
import json
import math
import os

import numpy as np
import pytest

from ance_b200 import evaluation as ev


def test_msmarco_mrr_matches_reference_golden(golden_dir):
    for case in json.load(open(os.path.join(golden_dir, "msmarco_mrr.json"))):
        rel = {int(k): v for k, v in case["relevant"].items()}
        ranked = {int(k): v for k, v in case["ranked"].items()}
        assert ev.msmarco_mrr(rel, ranked) == pytest.approx(case["mrr10"], abs=1e-15)
    with pytest.raises(IOError):
        ev.msmarco_mrr({1: [2]}, {3: [4]})


def _notebook_eval(q2id, p2id, pos, I, topN):
    """Literal restatement of the notebook's EvalDevQuery loop (cell 8) + brute-force trec measures."""
    prediction, ranked1000 = {}, {}
    total = labeled = Atotal = Alabeled = 0
    for qi in range(len(I)):
        seen, qid = set(), int(q2id[qi])
        prediction[qid] = {}
        ranked1000.setdefault(qid, [0] * 1000)
        rank = 0
        for idx in list(I[qi])[:topN]:
            pid = int(p2id[idx])
            if pid not in seen:
                ranked1000[qid][rank] = pid
                Atotal += 1
                if pid not in pos.get(qid, {}):
                    Alabeled += 1
                if rank < 10:
                    total += 1
                    if pid not in pos.get(qid, {}):
                        labeled += 1
                rank += 1
                prediction[qid][pid] = -rank
                seen.add(pid)
    nd = mp = rr = rc = 0.0
    n = 0
    for qid, run in prediction.items():
        if qid not in pos:
            continue
        n += 1
        order = [p for p, _ in sorted(run.items(), key=lambda kv: -kv[1])]
        rels = [pos[qid].get(p, 0) for p in order]
        num_rel = sum(1 for v in pos[qid].values() if v > 0)
        dcg = sum(g / math.log2(i + 2) for i, g in enumerate(rels[:10]) if g > 0)
        ideal = sorted((v for v in pos[qid].values() if v > 0), reverse=True)[:10]
        idcg = sum(g / math.log2(i + 2) for i, g in enumerate(ideal))
        nd += dcg / idcg if idcg else 0.0
        hit = 0
        ap = 0.0
        for i, g in enumerate(rels[:10]):
            if g > 0:
                hit += 1
                ap += hit / (i + 1)
        mp += ap / num_rel if num_rel else 0.0
        first = next((i for i, g in enumerate(rels) if g > 0), None)
        rr += 1.0 / (first + 1) if first is not None else 0.0
        rc += sum(1 for g in rels if g > 0) / num_rel if num_rel else 0.0
    rel = {int(q): [p for p in d if p > 0] for q, d in pos.items()}
    return {"ndcg@10": nd / n, "eval_query_cnt": n, "map@10": mp / n, "mrr": rr / n, "recall@%d" % topN: rc / n,
            "hole_rate@10": labeled / total, "hole_rate": Alabeled / Atotal, "ms_mrr@10": ev.msmarco_mrr(rel, ranked1000)}


@pytest.mark.parametrize("chunks", [1, 3])
def test_eval_dev_query_full_equals_notebook_loop(chunks):
    rng = np.random.default_rng(5 + chunks)
    n_docs, n_q, topN = 80, 30, 25
    p2id = np.repeat(np.arange(1, n_docs + 1), chunks)            # several vectors per document (MaxP) when chunks > 1
    rng.shuffle(p2id)
    q2id = rng.permutation(200)[:n_q]
    pos = {int(q): {int(p): int(rng.integers(0, 3)) for p in rng.choice(np.arange(1, n_docs + 1), size=3, replace=False)}
           for q in q2id[:24]}                                    # 6 queries without judgements; some labels are 0
    I = np.stack([rng.permutation(len(p2id))[:40] for _ in range(n_q)])
    for q in range(0, n_q, 2):                                    # plant a relevant document early for half the queries
        if int(q2id[q]) in pos:
            want = next(iter(pos[int(q2id[q])]))
            I[q, int(rng.integers(0, 12))] = int(np.where(p2id == want)[0][0])
    got = ev.eval_dev_query_full(q2id, p2id, pos, I, topN)
    want = _notebook_eval(q2id, p2id, pos, I, topN)
    assert got.keys() == want.keys()
    for k in want:
        assert got[k] == pytest.approx(want[k], abs=1e-12), k
    with pytest.raises(IndexError):
        ev.eval_dev_query_full(q2id, p2id, pos, -np.ones((n_q, 5), dtype=np.int64), 5)


def test_rerank_ranks_only_the_first_stage_candidates_and_load_dumps_roundtrip(tmp_path):
    rng = np.random.default_rng(9)
    P = rng.standard_normal((60, 16)).astype(np.float32)
    Q = rng.standard_normal((5, 16)).astype(np.float32)
    p2id = np.repeat(np.arange(100, 130), 2)                      # two rows per pid
    q2id = np.arange(5)
    first = {0: [100, 105, 129], 1: [111], 2: [], 3: [100, 999], 4: list(range(100, 130))}
    out = ev.rerank(Q, q2id, P, p2id, first)
    for q in range(5):
        rows = [r for pid in first[q] for r in np.where(p2id == pid)[0]]
        assert sorted(out[q].tolist()) == sorted(rows)
        s = (P[out[q]].astype(np.float64) @ Q[q].astype(np.float64)).astype(np.float32)
        assert (np.diff(s) <= 0).all()                            # descending exact scores
    # dumps under the reference's names, two ranks
    for r, sl in enumerate((slice(0, 35), slice(35, 60))):
        np.save(tmp_path / f"passage_7__emb_p__data_obj_{r}.npy", P[sl])
        np.save(tmp_path / f"passage_7__embid_p__data_obj_{r}.npy", p2id[sl])
    e, i = ev.load_dumps(str(tmp_path), "passage_7_")
    assert np.array_equal(e, P) and np.array_equal(i, p2id)
    with pytest.raises(FileNotFoundError):
        ev.load_dumps(str(tmp_path), "dev_query_7_")
