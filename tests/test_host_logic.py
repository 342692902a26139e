"""Host-side logic of the product (token-cache I/O, row layout, post-processing, bookkeeping) against
the oracle and the reference-generated golden files.  CPU only, no compute kernels."""
import argparse
import json
import os
import random

import numpy as np
import pytest
import torch

from ance_b200 import postprocess
from ance_b200.data import (EmbeddingCache, GetProcessingFn, GetProcessingFnDPR, StreamingDataset,
                            StridedBatchReader)
from ance_b200.drivers import run_ann_data_gen as drv
from oracle import flat_ip_oracle, refresh_oracle
from tests.test_oracle_golden import _postprocess_inputs


def _cache(tmp_path, n=37, L=16, seed=1, name="passages"):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, L + 1, size=n)
    ids = np.full((n, L), 1, dtype=np.int32)
    for i in range(n):
        ids[i, :lens[i]] = rng.integers(3, 1000, size=lens[i])
    base = str(tmp_path / name)
    refresh_oracle.write_cache(base, lens, ids)
    return base, lens, ids


def test_embedding_cache_contract(tmp_path, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "refresh_io.json")))
    ids = np.array(g["ids"], dtype=np.int32)
    lens = np.array(g["lens"])
    base = str(tmp_path / "passages")
    refresh_oracle.write_cache(base, lens, ids)
    cache = EmbeddingCache(base)
    assert len(cache) == g["N"] and cache.record_size == 4 + 4 * g["L"] and cache.dtype == np.int32
    with cache as c:
        for i in (0, 5, g["N"] - 1):
            l, p = c[i]
            assert l == lens[i] and (p == ids[i]).all() and p.dtype == np.int32
        with pytest.raises(IndexError):
            c[-1]
        with pytest.raises(IndexError):
            c[g["N"]]
        assert [l for l, _ in c] == lens.tolist()
        args = argparse.Namespace(max_seq_length=g["L"], max_query_length=g["L"])
        rec = GetProcessingFn(args, query=False)(c[5], 5)[0]
        assert rec[0].tolist() == g["proc_fn"]["ids"] and rec[1].int().tolist() == g["proc_fn"]["mask"]
        assert rec[2].tolist() == g["proc_fn"]["type"] and int(rec[3]) == g["proc_fn"]["idx"]
        assert [str(t.dtype) for t in rec] == g["proc_fn"]["dtypes"]
        assert GetProcessingFn(args, query=True)(c[5], 5)[0][2].tolist() == g["proc_fn_query_type"]
        ds = StreamingDataset(c, GetProcessingFn(args, query=False))
        assert [int(r[3]) for r in ds] == g["stream_idx_w1"]
    mm = cache.memmap()
    assert (mm["len"] == lens).all() and (mm["ids"] == ids).all()


def test_processing_fn_dpr():
    ids = np.array([101, 7, 9, 102, 0, 0], dtype=np.int32)
    rec = GetProcessingFnDPR(None)((4, ids), 3)[0]
    exp = refresh_oracle.processing_fn_dpr(ids, 3)
    assert rec[0].tolist() == exp[0].tolist() and rec[1].tolist() == exp[1].tolist() and int(rec[3]) == 3


@pytest.mark.parametrize("W", [1, 2, 4, 8])
@pytest.mark.parametrize("B", [4, 16, 100])
def test_strided_reader_matches_reference_batches(tmp_path, W, B):
    base, lens, ids = _cache(tmp_path)
    cache = EmbeddingCache(base)
    for rank in range(W):
        want = refresh_oracle.rank_records(37, W, rank)
        got_idx, got_ids, got_lens = [], [], []
        batches = 0
        for bi, bl, bx in StridedBatchReader(cache, B, rank=rank, world_size=W, pin=False):
            assert bi.dtype == torch.int32 and bl.dtype == torch.int32 and bx.dtype == torch.int64
            assert bi.shape[0] <= B
            got_idx += bx.tolist()
            got_ids.append(bi.numpy())
            got_lens += bl.tolist()
            batches += 1
        assert got_idx == want
        assert batches == len(StridedBatchReader(cache, B, rank=rank, world_size=W, pin=False))
        if want:
            assert (np.concatenate(got_ids) == ids[want]).all() and got_lens == lens[want].tolist()


def test_strided_reader_empty_and_len_mismatch(tmp_path):
    base = str(tmp_path / "empty")
    refresh_oracle.write_cache(base, np.zeros(0, dtype=np.int32), np.zeros((0, 8), dtype=np.int32))
    assert list(StridedBatchReader(EmbeddingCache(base), 4, pin=False)) == []
    base, _, _ = _cache(tmp_path, name="p2")
    with pytest.raises(ValueError):
        StridedBatchReader(EmbeddingCache(base), 4, max_len=128, pin=False)


@pytest.mark.parametrize("n,B,C", [(10, 4, 4), (7, 16, 4), (16, 4, 2), (5, 1, 3)])
def test_rows_from_batches_is_chunk_major(n, B, C):
    idx = np.arange(100, 100 + n, dtype=np.int64)
    emb = torch.arange(n * C * 2, dtype=torch.float32).reshape(n, C, 2)
    rows, ids = drv.rows_from_batches(emb, idx, B)
    # the reference: for each batch, for each chunk, the rows of all docs of the batch (run_ann_data_gen.py:183-186)
    want_rows, want_ids = [], []
    for b0 in range(0, n, B):
        for c in range(C):
            want_rows.append(emb[b0:b0 + B, c])
            want_ids += idx[b0:b0 + B].tolist()
    assert ids.tolist() == want_ids and torch.equal(rows, torch.cat(want_rows))
    assert ids.tolist() == [100 + i for i in refresh_oracle.rank_embedding2id(n, 1, 0, B, chunks=C)]


@pytest.mark.parametrize("variant", ["shuffle", "topk"])
def test_postprocess_reference_sampler_is_byte_identical(golden_dir, tmp_path, variant):
    g = json.load(open(os.path.join(golden_dir, "refresh_postprocess.json")))
    Prow, Qrow, Drow, p2id, q2id, d2id, train_pos, dev_pos = _postprocess_inputs(g)
    v = g[variant]
    _, dev_I = flat_ip_oracle.search_bruteforce(Prow, Drow, 100)
    ndcg, n = postprocess.eval_dev_query(d2id, p2id, dev_pos, dev_I)
    assert n == v["num_queries_dev"] and ndcg == pytest.approx(v["ndcg"], abs=1e-12)
    s, e = postprocess.query_chunk(len(Qrow), v["output_num"], v["chunk_factor"])
    assert (s, e) == refresh_oracle.query_chunk(len(Qrow), v["output_num"], v["chunk_factor"])
    _, I = flat_ip_oracle.search_bruteforce(Prow, Qrow[s:e], 20)
    random.seed(0)
    negs, mrr, nq = postprocess.generate_negatives(q2id[s:e], p2id, train_pos, I, 5, select_topk=v["topk_mrr"],
                                                   sampler="reference")
    orc_negs, orc_mrr, orc_nq = refresh_oracle.generate_negatives(
        q2id[s:e], p2id, train_pos, I, set(q2id[s:e].tolist()), 5, v["topk_mrr"], random.Random(0))
    assert negs == orc_negs and nq == orc_nq and mrr == pytest.approx(orc_mrr, abs=1e-12)
    path = str(tmp_path / "ann_training_data")
    postprocess.write_training_data(path, q2id[s:e], train_pos, negs, sampler="reference")
    assert open(path).read() == v["training_data"]
    postprocess.write_ndcg(str(tmp_path / "ann_ndcg"), ndcg, "ckpt/checkpoint-7/")
    assert json.load(open(tmp_path / "ann_ndcg")) == json.loads(v["ndcg_file"])


def test_postprocess_fast_sampler_invariants():
    rng = np.random.default_rng(0)
    nq, k, n_p = 50, 40, 300
    p2id = rng.integers(0, 120, size=n_p)  # many rows share a pid (MaxP-like duplicates)
    I = np.stack([rng.permutation(n_p)[:k] for _ in range(nq)])
    q2id = np.arange(nq)
    pos = {q: int(p2id[I[q, 3]]) for q in range(nq)}  # the positive is among the neighbours
    negs, _, n = postprocess.generate_negatives(q2id, p2id, pos, I, 7, sampler="fast", seed=3)
    assert n == nq
    for q in range(nq):
        cand = set(p2id[I[q]].tolist()) - {pos[q]}
        assert len(negs[q]) == min(7, len(cand)) and len(set(negs[q])) == len(negs[q])
        assert set(negs[q]) <= cand
    again, _, _ = postprocess.generate_negatives(q2id, p2id, pos, I, 7, sampler="fast", seed=3)
    assert again == negs


def test_mrr_and_break_semantics_match_oracle():
    rng = np.random.default_rng(1)
    for trial in range(20):
        nq, k, n_p = 8, 12, 60
        p2id = rng.integers(0, 25, size=n_p)
        I = np.stack([rng.permutation(n_p)[:k] for _ in range(nq)])
        q2id = np.arange(nq)
        pos = {q: int(p2id[I[q, rng.integers(0, k)]]) for q in range(nq)}
        for topk in (False, True):
            random.seed(trial)
            a = postprocess.generate_negatives(q2id, p2id, pos, I, 3, select_topk=topk, sampler="reference")
            b = refresh_oracle.generate_negatives(q2id, p2id, pos, I, set(range(nq)), 3, topk, random.Random(trial))
            assert a[0] == b[0] and a[1] == pytest.approx(b[1], abs=1e-12) and a[2] == b[2]


def test_ndcg_matches_oracle_graded():
    rng = np.random.default_rng(2)
    for _ in range(50):
        ranked = rng.permutation(30)[:12].tolist()
        qrel = {int(p): int(rng.integers(0, 4)) for p in rng.permutation(30)[:8]}
        assert postprocess.ndcg_cut(ranked, qrel) == pytest.approx(refresh_oracle.ndcg_cut(ranked, qrel), abs=1e-15)
    assert postprocess.ndcg_cut([1, 2], {}) == 0.0


def test_negative_label_refused():
    I = np.array([[0, -1]])
    with pytest.raises(IndexError):
        postprocess.generate_negatives(np.array([0]), np.array([5, 6]), {0: 5}, I, 1, sampler="fast", seed=0)


def test_bookkeeping(tmp_path, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "refresh_postprocess.json")))
    for p, n in g["checkpoint_no"].items():
        assert drv.get_checkpoint_no(p) == n
    out = tmp_path / "out"
    assert drv.get_latest_ann_data(str(out)) == (-1, None, None)
    out.mkdir()
    assert drv.get_latest_ann_data(str(out)) == (-1, None, None)
    for n in (0, 4, 11):
        postprocess.write_ndcg(str(out / f"ann_ndcg_{n}"), 0.1 * n, f"c{n}")
    no, path, js = drv.get_latest_ann_data(str(out))
    assert no == 11 and path.endswith("ann_training_data_11") and js["checkpoint"] == "c11"
    assert (no, path) == refresh_oracle.latest_ann_data(str(out))[:2]
    # a refresher killed between open() and os.replace() leaves its staged file behind: the UNMODIFIED trainer's scan
    # (utils/util.py:227-236: int(name[len('ann_ndcg_'):]) on every file with that prefix — restated literally in
    # refresh_oracle.latest_ann_data) must not trip over it
    for name in ("ann_ndcg_12", "ann_training_data_12"):
        staged = postprocess.staging_path(str(out / name))
        assert os.path.dirname(staged) == str(out) and not os.path.basename(staged).startswith("ann_")
        open(staged, "w").write("partial")
    assert refresh_oracle.latest_ann_data(str(out))[:2] == (11, path)
    assert drv.get_latest_ann_data(str(out))[:2] == (11, path)
    # checkpoints count only once scheduler.pt exists (run_ann_data_gen.py:60-63)
    tr = tmp_path / "train"
    args = argparse.Namespace(training_dir=str(tr), init_model_dir="init/")
    assert drv.get_latest_checkpoint(args) == ("init/", 0)
    (tr / "checkpoint-100").mkdir(parents=True)
    (tr / "checkpoint-200").mkdir()
    assert drv.get_latest_checkpoint(args) == ("init/", 0)
    (tr / "checkpoint-100" / "scheduler.pt").write_bytes(b"")
    assert drv.get_latest_checkpoint(args) == (os.path.join(str(tr), "checkpoint-100") + "/", 100)


def test_cli_has_every_reference_flag():
    ref_flags = ["--data_dir", "--training_dir", "--init_model_dir", "--last_checkpoint_dir", "--model_type",
                 "--output_dir", "--cache_dir", "--end_output_num", "--max_seq_length", "--max_query_length",
                 "--max_doc_character", "--per_gpu_eval_batch_size", "--ann_chunk_factor", "--topk_training",
                 "--negative_sample", "--ann_measure_topk_mrr", "--only_keep_latest_embedding_file", "--no_cuda",
                 "--local_rank", "--server_ip", "--server_port", "--inference", "--config_name", "--tokenizer_name"]
    a = drv.get_arguments(["--data_dir", "d", "--training_dir", "t", "--init_model_dir", "i", "--model_type",
                           "rdot_nll", "--output_dir", "o", "--cache_dir", "c"])
    for f in ref_flags:
        assert hasattr(a, f[2:]), f
    # reference defaults (run_ann_data_gen.py:443-627)
    assert (a.max_seq_length, a.max_query_length, a.per_gpu_eval_batch_size, a.ann_chunk_factor, a.topk_training,
            a.negative_sample, a.end_output_num, a.local_rank) == (128, 64, 128, 5, 500, 5, -1, -1)


def test_registry_surface():
    from ance_b200.models import MSMarcoConfigDict
    assert list(MSMarcoConfigDict) == ["rdot_nll", "rdot_nll_multi_chunk", "dpr", "seeddot_nll"]
    for name, cfg in MSMarcoConfigDict.items():
        assert cfg.name == name and cfg.use_mean is False
        for attr in ("model_class", "process_fn", "tokenizer_class", "config_class"):
            assert getattr(cfg, attr) is not None
    for name in ("rdot_nll", "rdot_nll_multi_chunk", "dpr"):
        cls = MSMarcoConfigDict[name].model_class
        assert hasattr(cls, "query_emb") and hasattr(cls, "body_emb")
    # seeddot_nll resolves and FALLS BACK to the reference's stock module (SURVEY.md par. 2.1 row 8): without the
    # reference checkout on sys.path that is a clear error, with it the reference's own class is what gets built
    import sys
    import types
    saved = {k: sys.modules.pop(k) for k in ("model", "model.models") if k in sys.modules}
    try:
        with pytest.raises(NotImplementedError, match="stock module"):
            MSMarcoConfigDict["seeddot_nll"].model_class()

        class RefSeed:                       # stand-in for /root/reference/model/models.py:201-221
            def __init__(self, config=None, model_argobj=None):
                self.config = config

            @classmethod
            def from_pretrained(cls, path, **kw):
                return cls(config=("loaded", path))

        pkg, mod = types.ModuleType("model"), types.ModuleType("model.models")
        mod.SEEDEncoderDot_NLL_LN = RefSeed
        pkg.models = mod
        sys.modules["model"], sys.modules["model.models"] = pkg, mod
        m = MSMarcoConfigDict["seeddot_nll"].model_class("cfg")
        assert type(m) is RefSeed and m.config == "cfg"
        assert MSMarcoConfigDict["seeddot_nll"].model_class.from_pretrained("ckpt/").config == ("loaded", "ckpt/")
    finally:
        sys.modules.pop("model", None)
        sys.modules.pop("model.models", None)
        sys.modules.update(saved)


def test_models_refuse_cpu_tensors():
    from transformers import RobertaConfig
    from ance_b200._lib import AnceError
    from ance_b200.models import RobertaDot_NLL_LN
    cfg = RobertaConfig(vocab_size=100, hidden_size=256, num_hidden_layers=1, num_attention_heads=4,
                        intermediate_size=512, max_position_embeddings=66, type_vocab_size=1, pad_token_id=1)
    m = RobertaDot_NLL_LN(cfg)
    keys = set(m.state_dict())
    assert "roberta.embeddings.word_embeddings.weight" in keys and "embeddingHead.weight" in keys and "norm.bias" in keys
    assert "roberta.encoder.layer.0.attention.self.query.weight" in keys
    with pytest.raises(AnceError):  # no CPU fallback
        m.query_emb(torch.zeros(1, 8, dtype=torch.long), torch.ones(1, 8, dtype=torch.long))


def test_fast_sampler_head_of_permutation_and_fallback():
    """The "fast" sampler reads only the head of a random permutation; rows whose head holds too few valid candidates
    (MaxP: many rows per document, or the positive hit repeatedly) must fall back to the whole list."""
    from ance_b200 import postprocess
    rng = np.random.default_rng(0)
    nq, k, n_rows, ns = 300, 200, 5000, 20
    I = np.stack([rng.permutation(n_rows)[:k] for _ in range(nq)]).astype(np.int64)
    q2id = np.arange(nq, dtype=np.int64) + 1000
    # rows 0..99: one pid per row (plenty of candidates); rows 100..299: only 25 distinct pids among the 200 rows
    p2id_many = np.arange(n_rows, dtype=np.int64)
    p2id_few = (np.arange(n_rows, dtype=np.int64) % 25)
    for p2id, lo, hi in ((p2id_many, 0, 100), (p2id_few, 100, 300)):
        pos = {int(q2id[r]): int(p2id[I[r, 3]]) for r in range(lo, hi)}
        negs, mrr, n = postprocess.generate_negatives(q2id[lo:hi], p2id, pos, I[lo:hi], ns, False, sampler="fast", seed=1)
        assert n == hi - lo and list(negs.keys()) == q2id[lo:hi].tolist() and mrr >= 0.0
        for r in range(lo, hi):
            got = negs[int(q2id[r])]
            cand = set(p2id[I[r]].tolist()) - {pos[int(q2id[r])]}
            assert len(got) == min(ns, len(cand)) and len(set(got)) == len(got) and set(got) <= cand
    # different seeds draw different negatives; the same seed is reproducible
    pos = {int(q2id[r]): int(I[r, 0]) for r in range(100)}
    a = postprocess.generate_negatives(q2id[:100], p2id_many, pos, I[:100], ns, False, sampler="fast", seed=5)[0]
    b = postprocess.generate_negatives(q2id[:100], p2id_many, pos, I[:100], ns, False, sampler="fast", seed=5)[0]
    c = postprocess.generate_negatives(q2id[:100], p2id_many, pos, I[:100], ns, False, sampler="fast", seed=6)[0]
    assert a == b and a != c
    # every candidate position is reachable (uniform head): over many seeds the union of picks covers the list
    seen = set()
    for s in range(40):
        seen |= set(postprocess.generate_negatives(q2id[:1], p2id_many, pos, I[:1], ns, False, sampler="fast", seed=s)[0][1000])
    assert len(seen) > 150


def test_postprocess_equals_oracle_property():
    """Random shapes of the negative-sampling scan (duplicated pids as in MaxP, the positive anywhere or absent, k smaller
    than the request): the vectorised scan must reproduce the reference loop exactly -- negatives, their order, MRR."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(seed=st.integers(0, 2 ** 31 - 1), nq=st.integers(1, 9), k=st.integers(1, 40), n_docs=st.integers(2, 30),
           chunks=st.sampled_from([1, 1, 4]), ns=st.integers(1, 12), topk=st.booleans())
    def run(seed, nq, k, n_docs, chunks, ns, topk):
        rng = np.random.default_rng(seed)
        n_rows = n_docs * chunks
        k2 = min(k, n_rows)
        I = np.stack([rng.permutation(n_rows)[:k2] for _ in range(nq)]).astype(np.int64)
        p2id = np.repeat(np.arange(n_docs, dtype=np.int64), chunks)[rng.permutation(n_rows)] if chunks > 1 \
            else rng.permutation(n_docs).astype(np.int64)
        q2id = rng.permutation(1000)[:nq].astype(np.int64)
        pos = {int(q): int(rng.integers(0, n_docs)) for q in q2id}
        random.seed(seed)
        got = postprocess.generate_negatives(q2id, p2id, pos, I, ns, select_topk=topk, sampler="reference")
        want = refresh_oracle.generate_negatives(q2id, p2id, pos, I, set(q2id.tolist()), ns, topk, random.Random(seed))
        assert got[0] == want[0] and got[2] == want[2] and got[1] == pytest.approx(want[1], abs=1e-12)

    run()


def test_array_negatives_and_native_writer_equal_the_python_path(tmp_path):
    """generate_negatives(as_arrays=True) + write_training_data_arrays (threaded row blocks, libance_b200's host writer)
    against the dict / Python-join path: same negatives, same bytes — including rows with fewer than `negative_sample`
    candidates and several row blocks."""
    from ance_b200 import postprocess
    rng = np.random.default_rng(12)
    nq, k, n_rows, ns = 700, 60, 4000, 9
    I = np.stack([rng.permutation(n_rows)[:k] for _ in range(nq)]).astype(np.int64)
    q2id = rng.permutation(5000)[:nq].astype(np.int64)
    p2id = np.arange(n_rows, dtype=np.int64)
    p2id[:3900] = np.arange(3900) % 7                     # most rows share 7 pids: many queries have < 9 distinct candidates
    pos = {int(q): int(p2id[I[r, 5]]) for r, q in enumerate(q2id)}
    old_chunk = postprocess._FAST_CHUNK
    postprocess._FAST_CHUNK = 128                          # several blocks -> the threaded path
    try:
        for select_topk in (False, True):
            d, mrr_d, _ = postprocess.generate_negatives(q2id, p2id, pos, I, ns, select_topk=select_topk, sampler="fast", seed=4)
            (mat, cnt), mrr_a, _ = postprocess.generate_negatives(q2id, p2id, pos, I, ns, select_topk=select_topk, sampler="fast",
                                                                  seed=4, as_arrays=True)
            assert mrr_d == mrr_a and (cnt <= ns).all() and (cnt < ns).any()
            assert d == {int(q): mat[r, :cnt[r]].tolist() for r, q in enumerate(q2id)}
            a, b = str(tmp_path / f"py{select_topk}"), str(tmp_path / f"native{select_topk}")
            n1 = postprocess.write_training_data(a, q2id, pos, d, sampler="fast", seed=4)
            n2 = postprocess.write_training_data_arrays(b, q2id, pos, mat, cnt, seed=4)
            assert n1 == n2 == nq and open(a, "rb").read() == open(b, "rb").read()
    finally:
        postprocess._FAST_CHUNK = old_chunk


def test_synthetic_data_dir_is_in_the_reference_formats(tmp_path):
    """ance_b200.synthetic (bench / full-refresh inputs): token caches readable through the EmbeddingCache contract, one
    positive per query in the qrels, and the multi-writer form (every rank of a job writes its share of the chunks into the
    pre-sized file) byte-identical to the single writer."""
    import filecmp
    from ance_b200 import synthetic
    from ance_b200.data import EmbeddingCache, StridedBatchReader
    from ance_b200.drivers.run_ann_data_gen import load_positive_ids
    d = tmp_path / "data"
    synthetic.write_marco_like_dir(str(d), 700, 60, 11, L_p=32, L_q=16, seed=3, vocab=500)
    with EmbeddingCache(str(d / "passages")) as c:
        assert len(c) == 700 and c.embedding_size == 32
        n, ids = c[699]
        assert 8 <= n <= 32 and ids[0] == 0 and ids[n - 1] == 2 and (ids[n:] == 1).all() and ids[:n].max() < 500
        with pytest.raises(IndexError):
            c[700]
    seen = [int(i) for _, _, idx in StridedBatchReader(EmbeddingCache(str(d / "train-query")), 7, 1, 3, pin=False) for i in idx]
    assert seen == list(range(1, 60, 3))
    train_pos, dev_pos = load_positive_ids(argparse.Namespace(data_dir=str(d)))
    assert sorted(train_pos) == list(range(60)) and sorted(dev_pos) == list(range(11))
    assert all(0 <= p < 700 for p in train_pos.values())
    # the same cache written by three "ranks" sharing the chunks
    synthetic.write_token_cache(str(tmp_path / "one"), 1000, 16, 9, 3, 4, 5, chunk=128)
    synthetic.presize_token_cache(str(tmp_path / "many"), 1000, 16)
    for part in range(3):
        synthetic.write_token_cache(str(tmp_path / "many"), 1000, 16, 9, 3, 4, 5, chunk=128, part=part, n_parts=3)
    assert filecmp.cmp(tmp_path / "one", tmp_path / "many", shallow=False)
    full = tmp_path / "full"
    synthetic.write_token_cache(str(full), 50, 16, 9, 3, 4, 5, full_length=True)
    assert (EmbeddingCache(str(full)).memmap()["len"] == 16).all()
