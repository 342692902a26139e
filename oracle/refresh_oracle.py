"""CPU oracle for the integer / byte side of the ANN refresh.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import this module.

Pure-Python restatements (small-case loops, deliberately slow and literal) of the reference's
  * token-cache record codec           utils/util.py:279-283 ; data/msmarco_data.py:252-258
  * GetProcessingFn                    data/msmarco_data.py:275-303 ; data/DPR_data.py:276-296
  * rank striding                      utils/util.py:318-329 (StreamingDataset.__iter__)
  * batched row layout incl. MaxP      drivers/run_ann_data_gen.py:167,183-192
  * gather order of barrier_array_merge utils/util.py:129-144
  * query chunking                     drivers/run_ann_data_gen.py:281-296
  * GenerateNegativePassaageID         drivers/run_ann_data_gen.py:339-396
  * EvalDevQuery                       drivers/run_ann_data_gen.py:399-440  (+ trec_eval's ndcg_cut_10,
       which the reference gets from the un-vendored, unpinned pytrec_eval: restated from the
       trec_eval definition, parity unpinned for that one function)
  * ann_training_data_N / ann_ndcg_N   drivers/run_ann_data_gen.py:315-334
  * get_latest_ann_data / get_checkpoint_no  utils/util.py:224-243
Pinned by tests/golden/refresh_*.json, produced by oracle/make_golden.py running the reference's
own functions (imported from /root/reference) on the same seeded inputs.
"""
from __future__ import annotations

import json
import math
import os
import random
import re
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------
# token cache
# ---------------------------------------------------------------------------------------------
def encode_record(length: int, ids: Sequence[int]) -> bytes:
    """4-byte BIG-endian length + native int32 token ids (msmarco_data.py:258 without the 8-byte id,
    which preprocess strips at msmarco_data.py:160-165)."""
    return int(length).to_bytes(4, "big") + np.asarray(ids, dtype=np.int32).tobytes()


def decode_record(buf: bytes) -> Tuple[int, np.ndarray]:
    """utils/util.py:279-283."""
    return int.from_bytes(buf[:4], "big"), np.frombuffer(buf[4:], dtype=np.int32)


def write_cache(base_path: str, lengths: Sequence[int], ids: np.ndarray) -> None:
    """Token cache + `_meta` exactly as preprocess leaves them (msmarco_data.py:160-176)."""
    n, L = ids.shape
    with open(base_path, "wb") as f:
        for i in range(n):
            f.write(encode_record(int(lengths[i]), ids[i]))
    with open(base_path + "_meta", "w") as f:
        json.dump({"type": "int32", "total_number": int(n), "embedding_size": int(L)}, f)


def processing_fn_marco(length: int, ids: np.ndarray, i: int, max_len: int, query: bool):
    """data/msmarco_data.py:275-303: (input_ids int32[L], attention_mask bool[L], token_type uint8[L], idx)."""
    pad = max(0, max_len - length)
    tt = ([0] if query else [1]) * length + [0] * pad
    mask = [1] * length + [0] * pad
    return (np.asarray(ids, dtype=np.int32), np.asarray(mask, dtype=bool), np.asarray(tt, dtype=np.uint8), int(i))


def processing_fn_dpr(ids: np.ndarray, i: int):
    """data/DPR_data.py:276-296: attention mask = ids != 0."""
    ids = np.asarray(ids, dtype=np.int32)
    return (ids, ids != 0, np.zeros_like(ids, dtype=np.uint8), int(i))


# ---------------------------------------------------------------------------------------------
# striding / layout
# ---------------------------------------------------------------------------------------------
def rank_records(n: int, world_size: int, rank: int) -> List[int]:
    """utils/util.py:325: record i belongs to rank i % world_size."""
    return [i for i in range(n) if i % world_size == rank]


def rank_embedding2id(n: int, world_size: int, rank: int, batch_size: int, chunks: int = 1) -> List[int]:
    """embedding2id a rank produces (run_ann_data_gen.py:167,183-192).  With chunks > 1 (MaxP) each
    batch contributes its ids once per chunk, chunk-major."""
    recs = rank_records(n, world_size, rank)
    out: List[int] = []
    for b in range(0, len(recs), batch_size):
        batch = recs[b:b + batch_size]
        for _ in range(chunks):
            out.extend(batch)
    return out


def merged_embedding2id(n: int, world_size: int, batch_size: int, chunks: int = 1) -> List[int]:
    """utils/util.py:129-144: rank 0 concatenates the per-rank arrays in rank order."""
    out: List[int] = []
    for r in range(world_size):
        out.extend(rank_embedding2id(n, world_size, r, batch_size, chunks))
    return out


def query_chunk(num_queries: int, output_num: int, chunk_factor: int) -> Tuple[int, int]:
    """run_ann_data_gen.py:281-296 (after the reference's `chunk_factor <= 0 -> 1` guard)."""
    effective = output_num % chunk_factor if chunk_factor > 0 else 0
    if chunk_factor <= 0:
        chunk_factor = 1
    per = num_queries // chunk_factor
    start = per * effective
    end = num_queries if effective == chunk_factor - 1 else start + per
    return start, end


# ---------------------------------------------------------------------------------------------
# post-processing
# ---------------------------------------------------------------------------------------------
def generate_negatives(query_embedding2id, passage_embedding2id, positives: Dict[int, int], I: np.ndarray,
                       effective_q_id: Iterable[int], negative_sample: int, select_topk: bool,
                       rng: random.Random):
    """run_ann_data_gen.py:339-396.  `rng` stands in for the module-level `random` the reference
    uses unseeded; with SelectTopK the first negative_sample+1 neighbours are taken in order,
    otherwise all k neighbours are visited in a shuffled order.  Returns (negatives, mrr_sum, n)."""
    eff = set(int(x) for x in effective_q_id)
    out: Dict[int, List[int]] = {}
    mrr = 0.0
    nq = 0
    for qi in range(I.shape[0]):
        qid = int(query_embedding2id[qi])
        if qid not in eff:
            continue
        nq += 1
        pos = positives[qid]
        row = I[qi].copy()
        if select_topk:
            sel = row[:negative_sample + 1]
        else:
            perm = list(range(I.shape[1]))
            rng.shuffle(perm)
            sel = row[perm]
        negs: List[int] = []
        cnt = 0
        rank = 0
        for idx in sel:
            pid = int(passage_embedding2id[idx])
            rank += 1
            if pid == pos:
                if rank <= 10:
                    mrr += 1.0 / rank
                continue
            if pid in negs:
                continue
            if cnt >= negative_sample:
                break
            negs.append(pid)
            cnt += 1
        out[qid] = negs
    return out, mrr, nq


def ndcg_cut(ranked_pids: Sequence[int], qrel: Dict[int, int], cut: int = 10) -> float:
    """trec_eval ndcg_cut_k: linear gain (= rel), log2(rank + 1) discount, ideal ranking from all
    judged documents of the query."""
    dcg = 0.0
    for r, pid in enumerate(ranked_pids[:cut], start=1):
        g = qrel.get(pid, 0)
        if g > 0:
            dcg += g / math.log2(r + 1)
    ideal = sorted((g for g in qrel.values() if g > 0), reverse=True)[:cut]
    idcg = sum(g / math.log2(r + 1) for r, g in enumerate(ideal, start=1))
    return dcg / idcg if idcg > 0 else 0.0


def eval_dev_query(query_embedding2id, passage_embedding2id, dev_qrels: Dict[int, Dict[int, int]], I: np.ndarray):
    """run_ann_data_gen.py:399-440: first 50 neighbours, pid-deduplicated ranking, mean ndcg_cut_10 over
    the queries that have both a run and a qrel (what pytrec_eval's evaluate() returns)."""
    pred: Dict[int, List[int]] = {}
    for qi in range(I.shape[0]):
        qid = int(query_embedding2id[qi])
        seen = set()
        ranked: List[int] = []
        for idx in I[qi, :50]:
            pid = int(passage_embedding2id[idx])
            if pid not in seen:
                ranked.append(pid)
                seen.add(pid)
        pred[qid] = ranked  # a later duplicate qid overwrites, as the dict assignment at :405 does
    total = 0.0
    n = 0
    for qid, ranked in pred.items():
        if qid not in dev_qrels:
            continue
        n += 1
        total += ndcg_cut(ranked, dev_qrels[qid], 10)
    return (total / n if n else 0.0), n


def training_data_lines(query_embedding2id, positives: Dict[int, int], negatives: Dict[int, List[int]],
                        effective_q_id: Iterable[int], rng: random.Random) -> List[str]:
    """run_ann_data_gen.py:318-329: shuffled query order, `qid \\t pos \\t n1,n2,...`."""
    eff = set(int(x) for x in effective_q_id)
    order = list(range(len(query_embedding2id)))
    rng.shuffle(order)
    lines = []
    for qi in order:
        qid = int(query_embedding2id[qi])
        if qid not in eff or qid not in positives:
            continue
        lines.append("{}\t{}\t{}\n".format(qid, positives[qid], ",".join(str(p) for p in negatives[qid])))
    return lines


def ndcg_json(ndcg: float, checkpoint: str) -> str:
    """run_ann_data_gen.py:331-334."""
    return json.dumps({"ndcg": ndcg, "checkpoint": checkpoint})


# ---------------------------------------------------------------------------------------------
# bookkeeping
# ---------------------------------------------------------------------------------------------
def checkpoint_no(path: str) -> int:
    """utils/util.py:224-226: the last run of digits in the path, 0 if none."""
    nums = re.findall(r"\d+", path)
    return int(nums[-1]) if nums else 0


def latest_ann_data(ann_dir: str):
    """utils/util.py:229-243."""
    prefix = "ann_ndcg_"
    if not os.path.exists(ann_dir):
        return -1, None, None
    nos = [int(f[len(prefix):]) for f in next(os.walk(ann_dir))[2] if f.startswith(prefix)]
    if not nos:
        return -1, None, None
    no = max(nos)
    with open(os.path.join(ann_dir, prefix + str(no))) as f:
        js = json.load(f)
    return no, os.path.join(ann_dir, "ann_training_data_" + str(no)), js


# ---------------------------------------------------------------------------------------------
# DPR post-processing (drivers/run_ann_data_gen_dpr.py:281-340, utils/dpr_utils.py:241-306)
# ---------------------------------------------------------------------------------------------
def dpr_words(text: str) -> List[str]:
    """SimpleTokenizer().tokenize(NFD(text)).words(uncased=True): utils/dpr_utils.py:253-257,267-306,331-338."""
    import unicodedata

    import regex
    pat = regex.compile(r"([\p{L}\p{N}\p{M}]+)|([^\p{Z}\p{C}])", flags=regex.IGNORECASE + regex.UNICODE + regex.MULTILINE)
    return [m.group().lower() for m in pat.finditer(unicodedata.normalize("NFD", text))]


def dpr_has_answer(answers: Sequence[str], text) -> bool:
    """utils/dpr_utils.py:241-264."""
    if text is None:
        return False
    words = dpr_words(text)
    for a in answers:
        aw = dpr_words(a)
        for i in range(0, len(words) - len(aw) + 1):
            if aw == words[i:i + len(aw)]:
                return True
    return False


def dpr_generate_negatives(passages, answers, query_embedding2id, passage_embedding2id, I, positives, negative_sample):
    """run_ann_data_gen_dpr.py:281-309: rank order, answer filter, neg_cnt advances on rejected candidates too."""
    out = {}
    for qi in range(I.shape[0]):
        qid = int(query_embedding2id[qi])
        negs, cnt = [], 0
        for pidx in I[qi]:
            doc = int(passage_embedding2id[pidx])
            if doc == positives[qid] or doc in negs:
                continue
            if cnt >= negative_sample:
                break
            if not dpr_has_answer(answers[qid], passages[doc][0]):
                negs.append(doc)
            cnt += 1
        out[qid] = negs
    return out


def dpr_validate(passages, answers, I, query_embedding2id, passage_embedding2id) -> List[float]:
    """run_ann_data_gen_dpr.py:312-340: hit@k for k = 1..n_docs."""
    n_docs = I.shape[1]
    hits = [0] * n_docs
    for qi in range(I.shape[0]):
        qid = int(query_embedding2id[qi])
        flags = [dpr_has_answer(answers[qid], passages[int(passage_embedding2id[p])][0]) for p in I[qi]]
        best = next((i for i, x in enumerate(flags) if x), None)
        if best is not None:
            for j in range(best, n_docs):
                hits[j] += 1
    return [v / I.shape[0] for v in hits]


# ======================================================================================================
# Trainer side of the refresh protocol (SURVEY.md par. 8(f) row 3): the consumers of ann_training_data_N
# ======================================================================================================
def parse_ann_line(line: str) -> Tuple[int, int, List[int]]:
    """data/msmarco_data.py:308-312 / 339-343."""
    a = line.split("\t")
    return int(a[0]), int(a[1]), [int(x) for x in a[2].split(",")]


def _padded(length: int, ids: np.ndarray, max_len: int, query: bool):
    """One record the way data/msmarco_data.py:275-303 presents it: ids as stored, mask 1 for the first `length`
    positions, token types 0 (query) / 1 (passage) on those positions."""
    pad = max(0, max_len - int(length))
    mask = [1] * int(length) + [0] * pad
    types = ([0] if query else [1]) * int(length) + [0] * pad
    return [int(x) for x in ids], mask, types


def training_pairs(lines, qlens, qids, plens, pids, max_query_length: int, max_seq_length: int):
    """data/msmarco_data.py:306-334: per negative a (query, positive, 1) and a (query, negative, 0) record."""
    out = []
    for line in lines:
        q, pos, negs = parse_ann_line(line)
        qr = _padded(qlens[q], qids[q], max_query_length, True)
        pr = _padded(plens[pos], pids[pos], max_seq_length, False)
        for n in negs:
            nr = _padded(plens[n], pids[n], max_seq_length, False)
            out.append([*qr, *pr, 1])
            out.append([*qr, *nr, 0])
    return out


def training_triplets(lines, qlens, qids, plens, pids, max_query_length: int, max_seq_length: int):
    """data/msmarco_data.py:337-362: per negative one (query, positive, negative) record."""
    out = []
    for line in lines:
        q, pos, negs = parse_ann_line(line)
        qr = _padded(qlens[q], qids[q], max_query_length, True)
        pr = _padded(plens[pos], pids[pos], max_seq_length, False)
        for n in negs:
            out.append([*qr, *pr, *_padded(plens[n], pids[n], max_seq_length, False)])
    return out


def nll_triplet_loss(logits_pos: np.ndarray, logits_neg: np.ndarray) -> float:
    """model/models.py:79-84: mean over the batch of -log_softmax([s+, s-])[0] = softplus(s- - s+)."""
    d = np.asarray(logits_neg, dtype=np.float64) - np.asarray(logits_pos, dtype=np.float64)
    return float(np.mean(np.maximum(d, 0.0) + np.log1p(np.exp(-np.abs(d)))))


def dot_logits(q: np.ndarray, x: np.ndarray) -> np.ndarray:
    """model/models.py:79-80: (q * x).sum(-1)."""
    return (np.asarray(q, dtype=np.float64) * np.asarray(x, dtype=np.float64)).sum(-1)


def maxp_logits(q: np.ndarray, x_chunks: np.ndarray, first_token_mask: np.ndarray) -> np.ndarray:
    """model/models.py:108-130: best chunk score; a chunk whose first token is padding is pushed down by 9999."""
    s = np.einsum("bd,bcd->bc", np.asarray(q, dtype=np.float64), np.asarray(x_chunks, dtype=np.float64))
    return (s + (1 - np.asarray(first_token_mask, dtype=np.float64)) * -9999.0).max(-1)
