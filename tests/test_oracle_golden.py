"""The oracle against the golden vectors produced by the reference's own code (oracle/make_golden.py).
CPU only.  If these fail the oracle has drifted from the reference and no parity claim holds."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import flat_ip_oracle, refresh_oracle
from oracle.encoder_oracle import BiEncoderOracle, RobertaDotOracle, random_roberta_state_dict


# ------------------------------------------------------------------------------------------------ search
def test_search_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, "search_kat.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    P = rng.standard_normal((3000, 64)).astype(np.float32)
    P[1500:1510] = P[10:20]
    Q = rng.standard_normal((16, 64)).astype(np.float32)
    Q[0] = P[12] * 2
    D, I = flat_ip_oracle.search_bruteforce(P, Q, 20)
    assert (I == g["I"]).all() and (D == g["D"]).all()
    D2, I2 = flat_ip_oracle.search(P, Q, 20, slack=32, q_block=5, p_block=999)
    assert (I2 == g["I"]).all() and (D2 == g["D"]).all()
    # planted tie: rows 12 and 1502 are identical, the smaller row number comes first
    assert I[0, 0] == 12 and I[0, 1] == 1502 and D[0, 0] == D[0, 1]


def test_search_faiss_contract_edges():
    rng = np.random.default_rng(0)
    P = rng.standard_normal((5, 8)).astype(np.float32)
    Q = rng.standard_normal((3, 8)).astype(np.float32)
    D, I = flat_ip_oracle.search_bruteforce(P, Q, 8)  # fewer rows than k: -1 / lowest-float padding
    assert (I[:, 5:] == -1).all() and (D[:, 5:] == np.finfo(np.float32).min).all()
    assert (np.diff(D[:, :5], axis=1) <= 0).all()
    D, I = flat_ip_oracle.search_bruteforce(P[:0], Q, 4)
    assert (I == -1).all()
    D, I = flat_ip_oracle.search_bruteforce(P, Q[:0], 4)
    assert D.shape == (0, 4)


def test_merge_shards_equals_global():
    rng = np.random.default_rng(5)
    P = rng.standard_normal((999, 32)).astype(np.float32)
    Q = rng.standard_normal((17, 32)).astype(np.float32)
    W, k = 4, 10
    order = np.concatenate([np.arange(r, 999, W) for r in range(W)])  # rank-major merged order
    Pm = P[order]
    Dg, Ig = flat_ip_oracle.search_bruteforce(Pm, Q, k)
    Ds, Is, off = [], [], 0
    for r in range(W):
        n = len(range(r, 999, W))
        d, i = flat_ip_oracle.search_bruteforce(Pm[off:off + n], Q, k)
        Ds.append(d)
        Is.append(np.where(i >= 0, i + off, -1))
        off += n
    Dm, Im = flat_ip_oracle.merge_shards(Ds, Is, k)
    assert (Im == Ig).all() and (Dm == Dg).all()


# ------------------------------------------------------------------------------------------------ I/O
def test_io_golden(golden_dir, tmp_path):
    g = json.load(open(os.path.join(golden_dir, "refresh_io.json")))
    ids = np.array(g["ids"], dtype=np.int32)
    lens = np.array(g["lens"])
    base = str(tmp_path / "passages")
    refresh_oracle.write_cache(base, lens, ids)
    raw = open(base, "rb").read()
    assert len(raw) == g["file_sha_len"] and raw[:4 + 4 * g["L"]].hex() == g["first_record_hex"]
    l, p = refresh_oracle.decode_record(raw[5 * (4 + 4 * g["L"]):6 * (4 + 4 * g["L"])])
    rec = refresh_oracle.processing_fn_marco(l, p, 5, g["L"], query=False)
    assert rec[0].tolist() == g["proc_fn"]["ids"] and rec[1].astype(int).tolist() == g["proc_fn"]["mask"]
    assert rec[2].tolist() == g["proc_fn"]["type"] and rec[3] == g["proc_fn"]["idx"]
    assert refresh_oracle.processing_fn_marco(l, p, 5, g["L"], query=True)[2].tolist() == g["proc_fn_query_type"]
    assert refresh_oracle.rank_records(g["N"], 1, 0) == g["stream_idx_w1"]


def test_layout_worked_example():
    """N=10, W=2, B=4."""
    assert refresh_oracle.merged_embedding2id(10, 2, 4) == [0, 2, 4, 6, 8, 1, 3, 5, 7, 9]
    assert refresh_oracle.rank_embedding2id(10, 2, 0, 4, chunks=4) == [0, 2, 4, 6] * 4 + [8] * 4


# ------------------------------------------------------------------------------------------------ post-processing
def _postprocess_inputs(g):
    rng = np.random.default_rng(g["seed"])
    n_p, n_q, n_dev, dim, W, B = g["n_p"], g["n_q"], g["n_dev"], g["dim"], g["W"], g["B"]
    P = rng.standard_normal((n_p, dim)).astype(np.float32)
    Q = rng.standard_normal((n_q, dim)).astype(np.float32)
    Qd = rng.standard_normal((n_dev, dim)).astype(np.float32)
    p2id = np.array(refresh_oracle.merged_embedding2id(n_p, W, B), dtype=np.int64)
    q2id = np.array(refresh_oracle.merged_embedding2id(n_q, W, B), dtype=np.int64)
    d2id = np.array(refresh_oracle.merged_embedding2id(n_dev, W, B), dtype=np.int64)
    train_pos = {int(q): int(rng.integers(0, n_p)) for q in range(n_q)}
    for q in range(0, n_q, 3):
        Q[q] = P[train_pos[q]] * 3 + Q[q] * 0.1
    dev_pos = {}
    for q in range(n_dev):
        dev_pos[q] = {int(rng.integers(0, n_p)): 1}
        if q % 2 == 0:
            Qd[q] = P[next(iter(dev_pos[q]))] * 3 + Qd[q] * 0.1
    assert {str(k): v for k, v in train_pos.items()} == g["train_pos"]
    return P[p2id], Q[q2id], Qd[d2id], p2id, q2id, d2id, train_pos, dev_pos


@pytest.mark.parametrize("variant", ["shuffle", "topk"])
def test_postprocess_golden(golden_dir, variant):
    """oracle restatement == the reference's generate_new_ann output, byte for byte, under seed 0."""
    g = json.load(open(os.path.join(golden_dir, "refresh_postprocess.json")))
    Prow, Qrow, Drow, p2id, q2id, d2id, train_pos, dev_pos = _postprocess_inputs(g)
    v = g[variant]
    _, dev_I = flat_ip_oracle.search_bruteforce(Prow, Drow, 100)
    ndcg, n = refresh_oracle.eval_dev_query(d2id, p2id, dev_pos, dev_I)
    assert n == v["num_queries_dev"] and ndcg == pytest.approx(v["ndcg"], abs=1e-12)
    s, e = refresh_oracle.query_chunk(len(Qrow), v["output_num"], v["chunk_factor"])
    _, I = flat_ip_oracle.search_bruteforce(Prow, Qrow[s:e], 20)
    rng = random.Random(0)
    negs, _, _ = refresh_oracle.generate_negatives(q2id[s:e], p2id, train_pos, I, set(q2id[s:e].tolist()), 5,
                                                   v["topk_mrr"], rng)
    lines = refresh_oracle.training_data_lines(q2id[s:e], train_pos, negs, set(q2id[s:e].tolist()), rng)
    assert "".join(lines) == v["training_data"]
    assert json.loads(refresh_oracle.ndcg_json(ndcg, "ckpt/checkpoint-7/")) == json.loads(v["ndcg_file"])


def test_bookkeeping_golden(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "refresh_postprocess.json")))
    for p, n in g["checkpoint_no"].items():
        assert refresh_oracle.checkpoint_no(p) == n


# ------------------------------------------------------------------------------------------------ encoders
def test_encoder_oracle_rdot_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_rdot_nll.npz"))
    orc = RobertaDotOracle(random_roberta_state_dict(seed=int(g["seed"])))
    mask = np.arange(128)[None, :] < g["lens"][:, None]
    emb = orc.body_emb(torch.from_numpy(g["ids"]), torch.from_numpy(mask)).numpy()
    assert np.abs(emb - g["emb"]).max() < 2e-4
    qmask = np.arange(64)[None, :] < g["qlens"][:, None]
    qemb = orc.query_emb(torch.from_numpy(g["qids"]), torch.from_numpy(qmask)).numpy()
    assert np.abs(qemb - g["qemb"]).max() < 2e-4


def test_encoder_oracle_multi_chunk_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_multi_chunk.npz"))
    orc = RobertaDotOracle(random_roberta_state_dict(seed=int(g["seed"])))
    mask = np.arange(2048)[None, :] < g["lens"][:, None]
    emb = orc.body_emb_multi_chunk(torch.from_numpy(g["ids"]), torch.from_numpy(mask)).numpy()
    real = g["real_chunk"]
    assert np.abs(emb - g["emb"])[real].max() < 2e-4
    # all-padding chunks: one constant vector (2.3.0 additive-mask semantics, see oracle header)
    assert (emb[1, 2] == emb[1, 3]).all() and np.abs(emb[1, 2] - g["allpad_oracle_2_3_0"]).max() < 1e-5
    assert np.isfinite(emb).all()


def test_encoder_oracle_dpr_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_dpr.npz"))
    sd = {**random_roberta_state_dict(seed=int(g["seed_q"]), vocab=30522, max_pos=512, head=False,
                                      prefix="question_model."),
          **random_roberta_state_dict(seed=int(g["seed_c"]), vocab=30522, max_pos=512, head=False,
                                      prefix="ctx_model.")}
    orc = BiEncoderOracle(sd)
    ids = torch.from_numpy(g["ids"])
    assert np.abs(orc.body_emb(ids, ids != 0).numpy() - g["body_emb"]).max() < 2e-4
    assert np.abs(orc.query_emb(ids, ids != 0).numpy() - g["query_emb"]).max() < 2e-4


def test_c_restatement_agrees_with_numpy_oracle(golden_dir):
    """Two independent restatements of the search contract (BLAS + numpy sort vs plain C loops + heap) agree bit for
    bit: on the committed known-answer fixture, with exact duplicates (ties by row), with fewer rows than k, and on a
    case too large for the numpy brute force's fp64 score matrix to be convenient."""
    g = np.load(os.path.join(golden_dir, "search_kat.npz"))
    rng = np.random.default_rng(int(g["seed"]))
    P = rng.standard_normal((3000, 64)).astype(np.float32)
    P[1500:1510] = P[10:20]
    Q = rng.standard_normal((16, 64)).astype(np.float32)
    Q[0] = P[12] * 2
    D, I = flat_ip_oracle.search_c(P, Q, 20)
    assert (I == g["I"]).all() and (D == g["D"]).all()
    rng = np.random.default_rng(99)
    P = rng.standard_normal((5000, 768)).astype(np.float32)
    P[2500:2600] = P[:100]
    Q = rng.standard_normal((40, 768)).astype(np.float32)
    for k in (1, 7, 200):
        Dc, Ic = flat_ip_oracle.search_c(P, Q, k)
        Db, Ib = flat_ip_oracle.search_bruteforce(P, Q, k)
        assert (Ic == Ib).all() and (Dc == Db).all()
    Dc, Ic = flat_ip_oracle.search_c(P[:5], Q, 9)
    Db, Ib = flat_ip_oracle.search_bruteforce(P[:5], Q, 9)
    assert (Ic == Ib).all() and (Dc == Db).all() and (Ic[:, 5:] == -1).all()
    De, Ie = flat_ip_oracle.search_c(P[:0], Q, 3)
    assert (Ie == -1).all() and (De == flat_ip_oracle.LOWEST).all()
    P = rng.standard_normal((120000, 768)).astype(np.float32)
    Dc, Ic = flat_ip_oracle.search_c(P, Q[:8], 100)
    Ds, Is = flat_ip_oracle.search(P, Q[:8], 100)          # the blocked sgemm + canonical rescoring oracle
    assert (Ic == Is).all() and (Dc == Ds).all()
