"""End-to-end DPR refresh (BASELINE config 5 in miniature) through the drop-in DPR driver on the GPU, checked
against the oracle pipeline fed with the same embeddings."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import flat_ip_oracle, refresh_oracle
from oracle.encoder_oracle import BiEncoderOracle, random_roberta_state_dict

pytestmark = pytest.mark.gpu
VOCAB, LAYERS = 1000, 2
WORDS = ["alpha", "beta", "gamma", "delta", "omega", "paris", "rome", "1969", "moon", "tower"]


def _world(tmp_path, n_p=1500, n_q=64, n_t=24, L=128):
    rng = np.random.default_rng(0)
    data = tmp_path / "data"
    data.mkdir()

    def cache(name, n):
        lens = rng.integers(8, L + 1, size=n)
        ids = np.zeros((n, L), dtype=np.int32)
        for i, m in enumerate(lens):
            ids[i, :m] = rng.integers(103, VOCAB, size=m)
            ids[i, 0], ids[i, m - 1] = 101, 102
        refresh_oracle.write_cache(str(data / name), lens, ids)
        return ids

    ids = {k: cache(k, n) for k, n in (("passages", n_p), ("train-query", n_q), ("test-query", n_t),
                                       ("trivia-test-query", n_t))}
    texts = {i: (" ".join(rng.choice(WORDS, size=10)), f"title {i}") for i in range(n_p)}
    with open(data / "pid2offset", "w") as f:
        for i in range(n_p):
            f.write(f"{1000 + i}\t{i}\n")
    corp = tmp_path / "corpus"
    corp.mkdir()
    with open(corp / "psgs_w100.tsv", "w") as f:
        f.write("id\ttext\ttitle\n")
        for i in range(n_p):
            f.write(f"{1000 + i}\t{texts[i][0]}\t{texts[i][1]}\n")
    train_pos = [int(rng.integers(0, n_p)) for _ in range(n_q)]
    train_ans = [[str(rng.choice(WORDS))] for _ in range(n_q)]
    with open(data / "train-ann", "w") as f:
        for q in range(n_q):
            f.write(f"{q}\t{train_pos[q]}\t{train_ans[q]!r}\n")
    tests = {}
    for name in ("nq-test.csv", "trivia-test.csv"):
        ans = [[str(rng.choice(WORDS)), str(rng.choice(WORDS)) + " " + str(rng.choice(WORDS))] for _ in range(n_t)]
        with open(corp / name, "w") as f:
            for a in ans:
                f.write(f"question text\t{a!r}\n")
        tests[name] = ans
    sd = {**random_roberta_state_dict(seed=1, n_layer=LAYERS, vocab=VOCAB, max_pos=512, head=False, prefix="question_model."),
          **random_roberta_state_dict(seed=2, n_layer=LAYERS, vocab=VOCAB, max_pos=512, head=False, prefix="ctx_model.")}
    ck = tmp_path / "init_ckpt"
    torch.save({"model_dict": sd, "optimizer_dict": {}, "scheduler_dict": {}, "offset": 0, "epoch": 0,
                "encoder_params": {}}, str(ck))
    return data, corp, ck, ids, texts, train_pos, train_ans, tests, sd


def test_dpr_refresh_end_to_end(tmp_path):
    from ance_b200.drivers import run_ann_data_gen as base
    from ance_b200.drivers import run_ann_data_gen_dpr as ddrv
    data, corp, ck, ids, texts, train_pos, train_ans, tests, sd = _world(tmp_path)
    out = tmp_path / "ann"
    argv = ["--data_dir", str(data), "--training_dir", str(tmp_path / "none"), "--init_model_dir", str(ck),
            "--model_type", "dpr", "--output_dir", str(out), "--cache_dir", str(tmp_path / "cache"),
            "--end_output_num", "0", "--max_seq_length", "128", "--per_gpu_eval_batch_size", "16",
            "--topk_training", "20", "--negative_sample", "5", "--passage_path", str(corp), "--test_qa_path", str(corp),
            "--trivia_test_qa_path", str(corp), "--seed", "0"]
    args = ddrv.get_arguments(argv)
    args.num_hidden_layers, args.vocab_size = LAYERS, VOCAB   # miniature BERT (the reference is fixed at bert-base)
    base.set_env(args)
    ddrv.ann_data_gen(args)
    js = json.load(open(out / "ann_ndcg_0"))
    assert set(js) == {"top20", "top100", "top20_trivia", "top100_trivia", "checkpoint"} and js["checkpoint"] == str(ck)
    # --- oracle pipeline downstream of the GPU embeddings
    model = ddrv.load_model(args, str(ck))
    be = base.B200Backend(args, model, mask_mode="nonzero")
    E = {k: be.encode(str(data / k), k != "passages") for k in ids}
    P, p2id = E["passages"][0].cpu().numpy(), E["passages"][1]
    for name, key, ans, k20, k100 in (("test-query", "nq-test.csv", None, "top20", "top100"),
                                      ("trivia-test-query", "trivia-test.csv", None, "top20_trivia", "top100_trivia")):
        _, I = flat_ip_oracle.search(P, E[name][0].cpu().numpy(), 100)
        hits = refresh_oracle.dpr_validate(texts, tests[key], I, E[name][1], p2id)
        assert js[k20] == hits[19] and js[k100] == hits[99]
    Q, q2id = E["train-query"][0].cpu().numpy(), E["train-query"][1]
    _, I = flat_ip_oracle.search(P, Q, 20)
    negs = refresh_oracle.dpr_generate_negatives(texts, train_ans, q2id, p2id, I, train_pos, 5)
    order = list(range(len(q2id)))
    random.Random(0).shuffle(order)
    want = "".join("{}\t{}\t{}\n".format(int(q2id[i]), train_pos[int(q2id[i])], ",".join(map(str, negs[int(q2id[i])])))
                   for i in order)
    assert open(out / "ann_training_data_0").read() == want
    # --- embeddings vs the fp32 oracle bi-encoder (CLS, no head): relative tolerance, outputs are not normalised
    orc = BiEncoderOracle(sd, n_layer=LAYERS)
    x = torch.from_numpy(ids["passages"][:32])
    ref = orc.body_emb(x, x != 0)
    cos = torch.nn.functional.cosine_similarity(torch.from_numpy(P[:32]), ref, dim=-1).min().item()
    assert cos >= 0.9995
