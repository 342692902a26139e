"""End-to-end refresh through the drop-in driver on the GPU (BASELINE config 1 in spirit: a small synthetic
MS MARCO slice), checked against the oracle pipeline fed with the same embeddings: identical top-k labels
=> byte-identical ann_training_data_N under --reference_sampling --seed 0."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import flat_ip_oracle, refresh_oracle
from oracle.encoder_oracle import RobertaDotOracle, random_roberta_state_dict

pytestmark = pytest.mark.gpu


def _make_world(tmp_path, n_p=3000, n_q=200, n_dev=50, vocab=2000, n_layer=2):
    from transformers import RobertaConfig
    rng = np.random.default_rng(0)
    data = tmp_path / "data"
    data.mkdir()

    def cache(name, n, L, mean, sd, lo):
        lens = np.clip(rng.normal(mean, sd, size=n).round().astype(int), lo, L)
        ids = np.full((n, L), 1, dtype=np.int32)
        for i, m in enumerate(lens):
            ids[i, :m] = rng.integers(3, vocab, size=m)
            ids[i, 0], ids[i, m - 1] = 0, 2
        refresh_oracle.write_cache(str(data / name), lens, ids)
        return lens, ids

    caches = {"passages": cache("passages", n_p, 128, 76, 28, 8), "train-query": cache("train-query", n_q, 64, 9, 3, 4),
              "dev-query": cache("dev-query", n_dev, 64, 9, 3, 4)}
    train_pos = {q: int(rng.integers(0, n_p)) for q in range(n_q)}
    dev_pos = {q: {int(rng.integers(0, n_p)): 1} for q in range(n_dev)}
    with open(data / "train-qrel.tsv", "w") as f:
        for q, p in train_pos.items():
            f.write(f"{q}\t{p}\t1\n")
    with open(data / "dev-qrel.tsv", "w") as f:
        for q, d in dev_pos.items():
            for p, r in d.items():
                f.write(f"{q}\t{p}\t{r}\n")
    ckpt = tmp_path / "init_model"
    ckpt.mkdir()
    cfg = RobertaConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=n_layer, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1, bos_token_id=0, eos_token_id=2)
    cfg.save_pretrained(str(ckpt))
    sd = random_roberta_state_dict(seed=5, n_layer=n_layer, vocab=vocab)
    sd["classifier.dense.weight"] = torch.zeros(768, 768)  # present in real checkpoints, unused by the path
    torch.save(sd, str(ckpt / "pytorch_model.bin"))
    return data, ckpt, caches, train_pos, dev_pos, sd, n_layer


def _argv(data, ckpt, out, tmp_path, extra=()):
    return ["--data_dir", str(data), "--training_dir", str(tmp_path / "no_training_dir_yet"), "--init_model_dir",
            str(ckpt), "--model_type", "rdot_nll", "--output_dir", str(out), "--cache_dir", str(tmp_path / "cache"),
            "--end_output_num", "0", "--max_seq_length", "128", "--max_query_length", "64",
            "--per_gpu_eval_batch_size", "16", "--topk_training", "20", "--negative_sample", "5",
            "--ann_chunk_factor", "1", "--reference_sampling", "--seed", "0", *extra]


def test_refresh_end_to_end(tmp_path):
    from ance_b200.drivers import run_ann_data_gen as drv
    data, ckpt, caches, train_pos, dev_pos, sd, n_layer = _make_world(tmp_path)
    out = tmp_path / "ann"
    drv.main(_argv(data, ckpt, out, tmp_path))
    text = open(out / "ann_training_data_0").read()
    ndcg = json.load(open(out / "ann_ndcg_0"))
    assert ndcg["checkpoint"] == str(ckpt) and 0.0 <= ndcg["ndcg"] <= 1.0
    lines = text.splitlines()
    assert len(lines) == 200
    for ln in lines:  # file grammar the trainer parses (data/msmarco_data.py:338-343)
        q, p, negs = ln.split("\t")
        assert train_pos[int(q)] == int(p)
        n = [int(x) for x in negs.split(",")]
        assert len(n) == 5 and len(set(n)) == 5 and int(p) not in n
    # --- the same refresh with the oracle downstream of the (GPU) embeddings
    args = drv.get_arguments(_argv(data, ckpt, out, tmp_path))
    drv.set_env(args)
    _, _, model = drv.load_model(args, str(ckpt))
    be = drv.B200Backend(args, model)
    emb = {k: be.encode(str(data / k), k != "passages") for k in ("dev-query", "passages", "train-query")}
    P, p2id = emb["passages"][0].cpu().numpy(), emb["passages"][1]
    Q, q2id = emb["train-query"][0].cpu().numpy(), emb["train-query"][1]
    Dq, d2id = emb["dev-query"][0].cpu().numpy(), emb["dev-query"][1]
    assert p2id.tolist() == list(range(3000)) and P.shape == (3000, 768)
    _, dev_I = flat_ip_oracle.search(P, Dq, 100)
    want_ndcg, _ = refresh_oracle.eval_dev_query(d2id, p2id, dev_pos, dev_I)
    assert ndcg["ndcg"] == pytest.approx(want_ndcg, abs=1e-12)
    _, I = flat_ip_oracle.search(P, Q, 20)
    rng = random.Random(0)
    negs, _, _ = refresh_oracle.generate_negatives(q2id, p2id, train_pos, I, set(q2id.tolist()), 5, False, rng)
    want = "".join(refresh_oracle.training_data_lines(q2id, train_pos, negs, set(q2id.tolist()), rng))
    assert text == want
    # --- and the embeddings themselves against the fp32 oracle encoder (tolerance: tests/test_gpu_encoder.py)
    lens, ids = caches["passages"]
    orc = RobertaDotOracle(sd, n_layer=n_layer)
    ref = orc.body_emb(torch.from_numpy(ids[:64]), torch.from_numpy(np.arange(128)[None, :] < lens[:64, None]))
    cos = torch.nn.functional.cosine_similarity(torch.from_numpy(P[:64]), ref, dim=-1).min().item()
    assert cos >= 0.9995
    # (the retrieval-overlap gate against the fp32-encoded corpus: tests/test_gpu_encoder.py::test_retrieval_overlap_at_200)
    # resume bookkeeping: the next run starts at output 1 and, with no new checkpoint, only sleeps
    assert drv.get_latest_ann_data(str(out))[0] == 0


def test_maxp_refresh_end_to_end(tmp_path):
    """BASELINE config 4 in miniature: rdot_nll_multi_chunk, documents of 4 x 512 tokens, one index row per chunk
    in the reference's chunk-major-per-batch order (run_ann_data_gen.py:183-186), pid de-duplication downstream."""
    from transformers import RobertaConfig
    from ance_b200.drivers import run_ann_data_gen as drv
    rng = np.random.default_rng(1)
    vocab, n_layer, n_d, n_q, n_dev = 2000, 2, 150, 40, 12
    data = tmp_path / "data"
    data.mkdir()
    dlens = np.clip(rng.lognormal(6.6, 0.8, size=n_d).astype(int), 20, 2048)   # many documents leave chunks all-pad
    dids = np.full((n_d, 2048), 1, dtype=np.int32)
    for i, m in enumerate(dlens):
        dids[i, :m] = rng.integers(3, vocab, size=m)
        dids[i, 0] = 0
    refresh_oracle.write_cache(str(data / "passages"), dlens, dids)
    for name, n in (("train-query", n_q), ("dev-query", n_dev)):
        lens = rng.integers(4, 20, size=n)
        ids = np.full((n, 64), 1, dtype=np.int32)
        for i, m in enumerate(lens):
            ids[i, :m] = rng.integers(3, vocab, size=m)
            ids[i, 0] = 0
        refresh_oracle.write_cache(str(data / name), lens, ids)
    train_pos = {q: int(rng.integers(0, n_d)) for q in range(n_q)}
    with open(data / "train-qrel.tsv", "w") as f:
        for q, p in train_pos.items():
            f.write(f"{q}\t{p}\t1\n")
    with open(data / "dev-qrel.tsv", "w") as f:
        for q in range(n_dev):
            f.write(f"{q}\t{int(rng.integers(0, n_d))}\t1\n")
    ckpt = tmp_path / "init_model"
    ckpt.mkdir()
    RobertaConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=n_layer, num_attention_heads=12,
                  intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                  pad_token_id=1, bos_token_id=0, eos_token_id=2).save_pretrained(str(ckpt))
    torch.save(random_roberta_state_dict(seed=6, n_layer=n_layer, vocab=vocab), str(ckpt / "pytorch_model.bin"))
    out = tmp_path / "ann"
    argv = ["--data_dir", str(data), "--training_dir", str(tmp_path / "none"), "--init_model_dir", str(ckpt),
            "--model_type", "rdot_nll_multi_chunk", "--output_dir", str(out), "--cache_dir", str(tmp_path / "c"),
            "--end_output_num", "0", "--max_seq_length", "2048", "--max_query_length", "64",
            "--per_gpu_eval_batch_size", "16", "--topk_training", "40", "--negative_sample", "5",
            "--ann_chunk_factor", "1", "--reference_sampling", "--seed", "0"]
    drv.main(argv)
    args = drv.get_arguments(argv)
    drv.set_env(args)
    _, _, model = drv.load_model(args, str(ckpt))
    be = drv.B200Backend(args, model)
    P, p2id = be.encode(str(data / "passages"), False)
    assert P.shape == (n_d * 4, 768)
    assert p2id.tolist() == refresh_oracle.rank_embedding2id(n_d, 1, 0, 16, chunks=4)   # chunk-major per batch of 16
    Q, q2id = be.encode(str(data / "train-query"), True)
    _, I = flat_ip_oracle.search(P.cpu().numpy(), Q.cpu().numpy(), 40)
    rng2 = random.Random(0)
    negs, _, _ = refresh_oracle.generate_negatives(q2id, p2id, train_pos, I, set(q2id.tolist()), 5, False, rng2)
    want = "".join(refresh_oracle.training_data_lines(q2id, train_pos, negs, set(q2id.tolist()), rng2))
    assert open(out / "ann_training_data_0").read() == want
    for ln in want.splitlines():   # de-duplicated document ids
        n = ln.split("\t")[2].split(",")
        assert len(set(n)) == len(n)


def test_inference_dumps(tmp_path):
    from ance_b200.drivers import run_ann_data_gen as drv
    data, ckpt, caches, *_ = _make_world(tmp_path, n_p=300, n_q=20, n_dev=10)
    out = tmp_path / "ann"
    drv.main(_argv(data, ckpt, out, tmp_path, extra=("--inference",)))
    # names of utils/util.py:108-113 as called from run_ann_data_gen.py:213-226 (step 0: init model)
    for prefix, n in (("dev_query_0_", 10), ("passage_0_", 300)):
        e = np.load(out / f"{prefix}_emb_p__data_obj_0.npy")
        i = np.load(out / f"{prefix}_embid_p__data_obj_0.npy")
        assert e.shape == (n, 768) and e.dtype == np.float32 and i.tolist() == list(range(n))
    assert not (out / "ann_training_data_0").exists()


def test_offline_evaluation_of_inference_dumps(tmp_path):
    """SURVEY.md 8(f) row 4, second half: the dumps of `--inference` -> ance_b200.evaluation (cells 9-13 of the reference's
    notebook): the GPU full-rank search must equal the oracle's on the dumped embeddings, the metrics must equal the
    notebook loop restated in tests/test_evaluation.py, and a rerank over first-stage candidates that contain the positive
    must find it."""
    from ance_b200 import evaluation as ev
    from ance_b200.drivers import run_ann_data_gen as drv
    from tests.test_evaluation import _notebook_eval
    data, ckpt, caches, train_pos, dev_pos, *_ = _make_world(tmp_path, n_p=1200, n_q=20, n_dev=40)
    out = tmp_path / "ann"
    drv.main(_argv(data, ckpt, out, tmp_path, extra=("--inference",)))
    q, q2id = ev.load_dumps(str(out), "dev_query_0_")
    p, p2id = ev.load_dumps(str(out), "passage_0_")
    assert q.shape == (40, 768) and p.shape == (1200, 768)
    I = ev.full_rank(q, p, 100)
    _, Io = flat_ip_oracle.search(np.ascontiguousarray(p), np.ascontiguousarray(q), 100)
    assert (I == Io).all()
    rng = np.random.default_rng(2)
    first = {int(qid): [int(x) for x in rng.permutation(1200)[:50]] + list(dev_pos[int(qid)]) for qid in q2id}
    res = ev.evaluate_dumps(str(out), 0, dev_pos, topN=100, first_stage=first)
    want = _notebook_eval(q2id, p2id, dev_pos, I, 100)
    for k, v in want.items():
        assert res["full_rank"][k] == pytest.approx(v, abs=1e-12), k
    rr = res["rerank"]
    assert rr["recall@100"] == 1.0 and rr["eval_query_cnt"] == 40 and 0.0 < rr["mrr"] <= 1.0   # the positive is a candidate
    assert set(res["full_rank"]) == set(rr)


def test_no_cuda_flag_is_refused(tmp_path):
    from ance_b200.drivers import run_ann_data_gen as drv
    args = drv.get_arguments(["--data_dir", "d", "--training_dir", "t", "--init_model_dir", "i", "--model_type",
                              "rdot_nll", "--output_dir", "o", "--cache_dir", "c", "--no_cuda"])
    with pytest.raises(RuntimeError):
        drv.set_env(args)
