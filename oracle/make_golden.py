"""Generate tests/golden/* by running the REFERENCE's own code (imported from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

What is pinned:
  encoder_*.npz   outputs of the reference classes model/models.py:137-199,223-259 (RobertaDot_NLL_LN,
                  RobertaDot_CLF_ANN_NLL_MultiChunk, BiEncoder) on seeded random weights
                  (oracle.encoder_oracle.random_roberta_state_dict — regenerated, not stored) and fixed
                  token ids.  The installed transformers is 5.5.0, not the pinned 2.3.0: configs need
                  return_dict=False (models.py:45 asserts a tuple), and 5.x masks with dtype-min instead
                  of the additive -10000, which changes ONLY all-padding sequences (a constant vector
                  either way).  Those rows are stored separately (`allpad_*`) and the oracle keeps the
                  2.3.0 semantics the reference was written against.
  refresh_*.json  the reference's own post-processing (GenerateNegativePassaageID, EvalDevQuery, the
                  ann_training_data writer inside generate_new_ann, drivers/run_ann_data_gen.py:231-440)
                  executed with the un-importable third-party pieces stubbed: faiss.IndexFlatIP by
                  oracle.flat_ip_oracle, pytrec_eval by oracle.refresh_oracle.ndcg_cut, the model by
                  seeded random embeddings.  Also utils/util.py EmbeddingCache / StreamingDataset /
                  barrier merge order and data/msmarco_data.py GetProcessingFn on a tiny cache.
"""
import argparse
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import flat_ip_oracle, refresh_oracle  # noqa: E402
from oracle.encoder_oracle import (BiEncoderOracle, RobertaDotOracle, random_roberta_state_dict)  # noqa: E402


def stub_third_party():
    sys.path.append(REF)
    for m in ("pytrec_eval", "faiss", "tensorboardX"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["tensorboardX"].SummaryWriter = object
    import transformers
    # drivers/run_ann_data_gen.py:18-25 imports AdamW (removed in transformers 5.x) and never uses it
    transformers.__dict__.setdefault("AdamW", torch.optim.AdamW)


def make_ids(rng, B, L, lens, pad, vocab, cls=0, sep=2):
    ids = np.full((B, L), pad, dtype=np.int32)
    for b in range(B):
        n = int(lens[b])
        if n > 0:
            ids[b, :n] = rng.integers(3, vocab, size=n)
            ids[b, 0] = cls
            ids[b, n - 1] = sep
    return ids


def golden_encoders():
    from transformers import BertConfig, RobertaConfig
    import model.models as M

    rng = np.random.default_rng(0)
    # ---------------- rdot_nll ----------------
    sd = random_roberta_state_dict(seed=0)
    cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1, bos_token_id=0, eos_token_id=2, num_labels=2, return_dict=False,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    ref = M.RobertaDot_NLL_LN(cfg).eval()
    missing, unexpected = ref.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("classifier.") or "pooler" in k or "position_ids" in k for k in missing), missing
    lens = np.array([128, 97, 64, 33, 17, 8, 128, 76], dtype=np.int32)
    ids = make_ids(rng, 8, 128, lens, 1, 50265)
    mask = (np.arange(128)[None, :] < lens[:, None])
    with torch.no_grad():
        emb = ref.body_emb(torch.from_numpy(ids).long(), torch.from_numpy(mask).long()).numpy()
    orc = RobertaDotOracle(sd)
    o = orc.body_emb(torch.from_numpy(ids), torch.from_numpy(mask)).numpy()
    print("rdot_nll: oracle vs reference max abs diff", np.abs(o - emb).max())
    assert np.abs(o - emb).max() < 2e-4
    qlens = np.array([64, 9, 12, 5], dtype=np.int32)
    qids = make_ids(rng, 4, 64, qlens, 1, 50265)
    qmask = (np.arange(64)[None, :] < qlens[:, None])
    with torch.no_grad():
        qemb = ref.query_emb(torch.from_numpy(qids).long(), torch.from_numpy(qmask).long()).numpy()
    assert np.abs(orc.query_emb(torch.from_numpy(qids), torch.from_numpy(qmask)).numpy() - qemb).max() < 2e-4
    np.savez_compressed(os.path.join(GOLD, "encoder_rdot_nll.npz"), seed=0, ids=ids, lens=lens, emb=emb,
                        qids=qids, qlens=qlens, qemb=qemb)

    # ---------------- rdot_nll_multi_chunk ----------------
    refm = M.RobertaDot_CLF_ANN_NLL_MultiChunk(cfg).eval()
    refm.load_state_dict(sd, strict=False)
    dlens = np.array([2048, 700], dtype=np.int32)  # doc 1: chunk 0 full, chunk 1 partial, chunks 2-3 all padding
    dids = make_ids(rng, 2, 2048, dlens, 1, 50265)
    dmask = (np.arange(2048)[None, :] < dlens[:, None])
    with torch.no_grad():
        demb = refm.body_emb(torch.from_numpy(dids).long(), torch.from_numpy(dmask).long()).numpy()
    od = orc.body_emb_multi_chunk(torch.from_numpy(dids), torch.from_numpy(dmask)).numpy()
    real = np.array([[1, 1, 1, 1], [1, 1, 0, 0]], dtype=bool)
    print("multi_chunk: oracle vs reference, chunks with tokens:", np.abs(od - demb)[real].max(),
          " all-pad chunks (transformers 5.x vs 2.3.0 mask semantics):", np.abs(od - demb)[~real].max())
    assert np.abs(od - demb)[real].max() < 2e-4
    assert np.abs(od[1, 2] - od[1, 3]).max() == 0.0 and np.abs(demb[1, 2] - demb[1, 3]).max() == 0.0
    np.savez_compressed(os.path.join(GOLD, "encoder_multi_chunk.npz"), seed=0, ids=dids, lens=dlens, emb=demb,
                        real_chunk=real, allpad_oracle_2_3_0=od[1, 2])

    # ---------------- dpr ----------------
    bcfg = BertConfig(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                      intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                      pad_token_id=0, return_dict=False, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sdq = random_roberta_state_dict(seed=1, vocab=30522, max_pos=512, head=False, prefix="question_model.")
    sdc = random_roberta_state_dict(seed=2, vocab=30522, max_pos=512, head=False, prefix="ctx_model.")
    sdd = {**sdq, **sdc}
    be = M.BiEncoder.__new__(M.BiEncoder)  # the constructor downloads bert-base-uncased (models.py:228-233)
    torch.nn.Module.__init__(be)
    be.question_model = M.HFBertEncoder(bcfg)
    be.ctx_model = M.HFBertEncoder(bcfg)
    missing, unexpected = be.load_state_dict(sdd, strict=False)
    assert not unexpected, unexpected
    assert all("pooler" in k or "position_ids" in k for k in missing), missing
    be.eval()
    plens = np.array([256, 120, 31, 256], dtype=np.int32)
    pids = make_ids(rng, 4, 256, plens, 0, 30522, cls=101, sep=102)
    with torch.no_grad():
        pemb = be.body_emb(torch.from_numpy(pids).long(), torch.from_numpy(pids != 0).long()).numpy()
        qemb2 = be.query_emb(torch.from_numpy(pids).long(), torch.from_numpy(pids != 0).long()).numpy()
    bo = BiEncoderOracle(sdd)
    d1 = np.abs(bo.body_emb(torch.from_numpy(pids), torch.from_numpy(pids != 0)).numpy() - pemb).max()
    d2 = np.abs(bo.query_emb(torch.from_numpy(pids), torch.from_numpy(pids != 0)).numpy() - qemb2).max()
    print("dpr: oracle vs reference", d1, d2)
    assert d1 < 2e-4 and d2 < 2e-4
    np.savez_compressed(os.path.join(GOLD, "encoder_dpr.npz"), seed_q=1, seed_c=2, ids=pids, lens=plens,
                        body_emb=pemb, query_emb=qemb2)


def golden_io():
    """utils/util.py + data/msmarco_data.py on a tiny cache."""
    from utils.util import EmbeddingCache, StreamingDataset
    from data.msmarco_data import GetProcessingFn

    rng = np.random.default_rng(1)
    N, L = 37, 16
    lens = rng.integers(1, L + 1, size=N)
    ids = make_ids(rng, N, L, lens, 1, 1000)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        base = os.path.join(td, "passages")
        refresh_oracle.write_cache(base, lens, ids)
        raw = open(base, "rb").read()
        out["file_sha_len"] = len(raw)
        out["first_record_hex"] = raw[:4 + 4 * L].hex()
        cache = EmbeddingCache(base)
        with cache as c:
            l5, p5 = c[5]
            assert l5 == lens[5] and (p5 == ids[5]).all()
            args = argparse.Namespace(max_seq_length=L, max_query_length=L)
            fn = GetProcessingFn(args, query=False)
            rec = fn((l5, p5), 5)[0]
            out["proc_fn"] = {"ids": rec[0].tolist(), "mask": rec[1].int().tolist(), "type": rec[2].tolist(),
                              "idx": int(rec[3]), "dtypes": [str(t.dtype) for t in rec]}
            fnq = GetProcessingFn(args, query=True)
            out["proc_fn_query_type"] = fnq((l5, p5), 5)[0][2].tolist()
            # non-distributed streaming order
            ds = StreamingDataset(c, fn)
            out["stream_idx_w1"] = [int(r[3]) for r in ds]
    out["N"], out["L"] = N, L
    out["lens"] = lens.tolist()
    out["ids"] = ids.tolist()
    with open(os.path.join(GOLD, "refresh_io.json"), "w") as f:
        json.dump(out, f)


def golden_trainer_records():
    """data/msmarco_data.py:306-362 (the trainer's consumers of ann_training_data_N) on tiny caches."""
    from utils.util import EmbeddingCache, StreamingDataset
    from data.msmarco_data import GetTrainingDataProcessingFn, GetTripletTrainingDataProcessingFn

    rng = np.random.default_rng(7)
    n_p, n_q, Lp, Lq = 23, 6, 12, 8
    plens, qlens = rng.integers(1, Lp + 1, size=n_p), rng.integers(1, Lq + 1, size=n_q)
    pids, qids = make_ids(rng, n_p, Lp, plens, 1, 500), make_ids(rng, n_q, Lq, qlens, 1, 500)
    lines = []
    for q in (4, 0, 5, 2, 1):
        cand = rng.permutation(n_p)[:4].tolist()
        lines.append(f"{q}\t{cand[0]}\t{','.join(map(str, cand[1:]))}\n")
    out = {"n_p": n_p, "n_q": n_q, "Lp": Lp, "Lq": Lq, "plens": plens.tolist(), "qlens": qlens.tolist(),
           "pids": pids.tolist(), "qids": qids.tolist(), "lines": lines}
    with tempfile.TemporaryDirectory() as td:
        refresh_oracle.write_cache(os.path.join(td, "passages"), plens, pids)
        refresh_oracle.write_cache(os.path.join(td, "train-query"), qlens, qids)
        args = argparse.Namespace(max_seq_length=Lp, max_query_length=Lq)
        with EmbeddingCache(os.path.join(td, "train-query")) as qc, EmbeddingCache(os.path.join(td, "passages")) as pc:
            for name, mk in (("pairs", GetTrainingDataProcessingFn), ("triplets", GetTripletTrainingDataProcessingFn)):
                recs = list(StreamingDataset(lines, mk(args, qc, pc)))
                out[name] = [[t.int().tolist() if t.dim() else int(t) for t in r] for r in recs]
                out[name + "_dtypes"] = [str(t.dtype) for t in recs[0]]
    with open(os.path.join(GOLD, "trainer_records.json"), "w") as f:
        json.dump(out, f)


def golden_trainer_losses():
    """model/models.py:58-134,253-266: the reference's own forward() on triplets built from the encoder fixtures."""
    from transformers import BertConfig, RobertaConfig
    import model.models as M

    out = {}
    g = np.load(os.path.join(GOLD, "encoder_rdot_nll.npz"))
    sd = random_roberta_state_dict(seed=0)
    cfg = RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                        intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                        pad_token_id=1, bos_token_id=0, eos_token_id=2, num_labels=2, return_dict=False,
                        hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    T = lambda a: torch.from_numpy(np.asarray(a)).long()   # noqa: E731
    ids, lens, qids, qlens = g["ids"], g["lens"], g["qids"], g["qlens"]
    mask = (np.arange(128)[None, :] < lens[:, None])
    qmask = (np.arange(64)[None, :] < qlens[:, None])
    ref = M.RobertaDot_NLL_LN(cfg).eval()
    ref.load_state_dict(sd, strict=False)
    with torch.no_grad():   # triplet i = (query i, passage i, passage 4 + i)
        out["rdot_nll_loss"] = float(ref(T(qids), T(qmask), T(ids[:4]), T(mask[:4]), T(ids[4:]), T(mask[4:]))[0])
        out["rdot_nll_query_passthrough_ok"] = bool(torch.equal(ref(T(qids), T(qmask)), ref.query_emb(T(qids), T(qmask))))
    gm = np.load(os.path.join(GOLD, "encoder_multi_chunk.npz"))
    dids, dlens = gm["ids"], gm["lens"]
    dmask = (np.arange(2048)[None, :] < dlens[:, None])
    refm = M.RobertaDot_CLF_ANN_NLL_MultiChunk(cfg).eval()
    refm.load_state_dict(sd, strict=False)
    with torch.no_grad():   # two triplets: (q0, d0, d1), (q1, d1, d0); d1 has two all-padding chunks (MaxP bias -9999)
        out["multi_chunk_loss"] = float(refm(T(qids[:2]), T(qmask[:2]), T(dids), T(dmask), T(dids[::-1].copy()),
                                             T(dmask[::-1].copy()))[0])
    gd = np.load(os.path.join(GOLD, "encoder_dpr.npz"))
    pids = gd["ids"]
    bcfg = BertConfig(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                      intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12,
                      pad_token_id=0, return_dict=False, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sdd = {**random_roberta_state_dict(seed=1, vocab=30522, max_pos=512, head=False, prefix="question_model."),
           **random_roberta_state_dict(seed=2, vocab=30522, max_pos=512, head=False, prefix="ctx_model.")}
    be = M.BiEncoder.__new__(M.BiEncoder)
    torch.nn.Module.__init__(be)
    be.question_model, be.ctx_model = M.HFBertEncoder(bcfg), M.HFBertEncoder(bcfg)
    be.load_state_dict(sdd, strict=False)
    be.eval()
    pm = pids != 0
    with torch.no_grad():   # (question = passage rows 0-1 through the question tower, ctx a = rows 0-1, ctx b = rows 2-3)
        out["dpr_loss"] = float(be(T(pids[:2]), T(pm[:2]), T(pids[:2]), T(pm[:2]), T(pids[2:]), T(pm[2:]))[0])
    print("trainer losses:", out)
    with open(os.path.join(GOLD, "trainer_losses.json"), "w") as f:
        json.dump(out, f)


def golden_postprocess():
    """Run the reference's generate_new_ann with the third-party pieces stubbed."""
    sys.modules["transformers"].__dict__.setdefault("AdamW", torch.optim.AdamW)
    import drivers.run_ann_data_gen as drv

    rng = np.random.default_rng(2)
    n_p, n_q, n_dev, dim, W, B = 2000, 120, 40, 32, 4, 16
    P = rng.standard_normal((n_p, dim)).astype(np.float32)
    Q = rng.standard_normal((n_q, dim)).astype(np.float32)
    Qd = rng.standard_normal((n_dev, dim)).astype(np.float32)
    p2id = np.array(refresh_oracle.merged_embedding2id(n_p, W, B), dtype=np.int64)
    q2id = np.array(refresh_oracle.merged_embedding2id(n_q, W, B), dtype=np.int64)
    d2id = np.array(refresh_oracle.merged_embedding2id(n_dev, W, B), dtype=np.int64)
    # the passage rows follow the merged order: row r holds the embedding of record p2id[r]
    Prow, Qrow, Drow = P[p2id], Q[q2id], Qd[d2id]
    train_pos = {int(q): int(rng.integers(0, n_p)) for q in range(n_q)}
    # plant the positive among the neighbours of some queries so the skip / MRR branches run
    for q in range(0, n_q, 3):
        Q[q] = P[train_pos[q]] * 3 + Q[q] * 0.1
    Qrow = Q[q2id]
    dev_pos = {}
    for q in range(n_dev):
        dev_pos[q] = {int(rng.integers(0, n_p)): 1}
        if q % 2 == 0:
            pid = next(iter(dev_pos[q]))
            Qd[q] = P[pid] * 3 + Qd[q] * 0.1
    Drow = Qd[d2id]

    class FakeIndex:
        def __init__(self, d):
            self.x = None

        def add(self, x):
            self.x = x

        def search(self, q, k):
            return flat_ip_oracle.search_bruteforce(self.x, q, k)

    drv.faiss.omp_set_num_threads = lambda n: None
    drv.faiss.IndexFlatIP = FakeIndex

    class FakeEvaluator:
        def __init__(self, qrel, measures):
            self.qrel = qrel

        def evaluate(self, run):
            res = {}
            for qid, docs in run.items():
                if qid not in self.qrel:
                    continue
                ranked = [int(d) for d, _ in sorted(docs.items(), key=lambda kv: -kv[1])]
                res[qid] = {"ndcg_cut_10": refresh_oracle.ndcg_cut(
                    ranked, {int(k): v for k, v in self.qrel[qid].items()}, 10)}
            return res

    drv.pytrec_eval.RelevanceEvaluator = FakeEvaluator
    calls = iter([(Drow, d2id), (Prow, p2id), (Qrow, q2id)])
    drv.StreamInferenceDoc = lambda *a, **k: next(calls)
    drv.load_model = lambda args, ckpt: (None, None, None)
    drv.is_first_worker = lambda: True
    drv.GetProcessingFn = lambda *a, **k: None
    golden = {"n_p": n_p, "n_q": n_q, "n_dev": n_dev, "dim": dim, "W": W, "B": B, "seed": 2}
    with tempfile.TemporaryDirectory() as td:
        data_dir = os.path.join(td, "data")
        os.makedirs(data_dir)
        for nm, n in (("dev-query", n_dev), ("passages", n_p), ("train-query", n_q)):
            refresh_oracle.write_cache(os.path.join(data_dir, nm), np.ones(n, dtype=np.int32),
                                       np.zeros((n, 4), dtype=np.int32))
        out_dir = os.path.join(td, "out")
        os.makedirs(out_dir)
        for variant, topk_mrr, chunk_factor, output_num in (("shuffle", False, 1, 0), ("topk", True, 3, 4)):
            calls = iter([(Drow, d2id), (Prow, p2id), (Qrow, q2id)])
            drv.StreamInferenceDoc = lambda *a, **k: next(calls)
            args = argparse.Namespace(data_dir=data_dir, output_dir=out_dir, topk_training=20, negative_sample=5,
                                      ann_chunk_factor=chunk_factor, ann_measure_topk_mrr=topk_mrr, inference=False,
                                      rank=0, max_seq_length=4, max_query_length=4)
            random.seed(0)
            ndcg, nq_dev = drv.generate_new_ann(args, output_num, "ckpt/checkpoint-7/", train_pos, dev_pos, 7)
            golden[variant] = {
                "output_num": output_num, "chunk_factor": chunk_factor, "topk_mrr": topk_mrr,
                "ndcg": ndcg, "num_queries_dev": nq_dev,
                "training_data": open(os.path.join(out_dir, f"ann_training_data_{output_num}")).read(),
                "ndcg_file": open(os.path.join(out_dir, f"ann_ndcg_{output_num}")).read(),
            }
        # resume bookkeeping (utils/util.py:224-243)
        from utils.util import get_checkpoint_no, get_latest_ann_data
        golden["latest_ann"] = list(get_latest_ann_data(out_dir)[:1])
        golden["checkpoint_no"] = {p: get_checkpoint_no(p) for p in
                                   ("checkpoint-150000", "a/b12/checkpoint-7/", "nodigits", "x9/y")}
    golden["train_pos"] = {str(k): v for k, v in train_pos.items()}
    golden["dev_pos"] = {str(k): {str(a): b for a, b in v.items()} for k, v in dev_pos.items()}
    with open(os.path.join(GOLD, "refresh_postprocess.json"), "w") as f:
        json.dump(golden, f)
    print("postprocess golden: ndcg", golden["shuffle"]["ndcg"], golden["topk"]["ndcg"],
          "lines", golden["shuffle"]["training_data"].count("\n"), golden["topk"]["training_data"].count("\n"))


def golden_search():
    """No reference arithmetic exists to run here (faiss absent): the search KAT pins the ORACLE itself
    (definition-level brute force) so later edits cannot drift, with planted duplicates and ties."""
    rng = np.random.default_rng(3)
    P = rng.standard_normal((3000, 64)).astype(np.float32)
    P[1500:1510] = P[10:20]           # exact duplicate rows -> exact ties
    Q = rng.standard_normal((16, 64)).astype(np.float32)
    Q[0] = P[12] * 2                   # its top-1 is a tie between rows 12 and 1502
    D, I = flat_ip_oracle.search_bruteforce(P, Q, 20)
    D2, I2 = flat_ip_oracle.search(P, Q, 20, slack=32, q_block=8, p_block=700)
    assert (I == I2).all() and (D == D2).all()
    np.savez_compressed(os.path.join(GOLD, "search_kat.npz"), seed=3, D=D, I=I)


def golden_dpr():
    """utils/dpr_utils.py has_answer + the DPR driver's validate / GenerateNegativePassaageID (reference code)."""
    sys.modules["transformers"].__dict__.setdefault("AdamW", torch.optim.AdamW)
    from utils.dpr_utils import SimpleTokenizer, has_answer
    import drivers.run_ann_data_gen_dpr as ddrv

    tok = SimpleTokenizer()
    texts = [
        "Paris is the capital and most populous city of France.",
        "The Eiffel Tower (/ˈaɪfəl/ EYE-fəl) was built in 1889; Gustave Eiffel's company designed it.",
        "Zürich — Switzerland's largest city — lies at the north-western tip of Lake Zürich.",
        "In 1969, Apollo 11 landed on the Moon. Neil Armstrong said: 'one small step'.",
        "",
        "U.S.A. and u.s.a are tokenised differently from USA",
        "naïve café déjà vu",
    ]
    answers = [["paris"], ["Gustave Eiffel"], ["zurich"], ["Zürich"], ["apollo 11", "Apollo 12"], ["one small step"],
               ["small  step"], ["1969,"], ["u.s.a"], ["USA"], ["cafe"], ["café"], [""], ["the Moon.", "nothing"],
               ["north-western"], ["EYE-fəl"]]
    table = [[bool(has_answer(a, t, tok)) for t in texts] for a in answers]
    rng = np.random.default_rng(4)
    n_p, n_q, k = 60, 12, 10
    words = ["alpha", "beta", "gamma", "delta", "omega", "paris", "rome", "1969", "moon", "tower"]
    passages = {i: (" ".join(rng.choice(words, size=12)), "t%d" % i) for i in range(n_p)}
    q_answers = [[str(rng.choice(words))] + ([str(rng.choice(words)) + " " + str(rng.choice(words))] if q % 3 == 0 else [])
                 for q in range(n_q)]
    p2id = rng.permutation(n_p).astype(np.int64)
    q2id = rng.permutation(n_q).astype(np.int64)
    I = np.stack([rng.permutation(n_p)[:k] for _ in range(n_q)])
    pos = [int(rng.integers(0, n_p)) for _ in range(n_q)]
    for q in range(0, n_q, 2):
        pos[int(q2id[q])] = int(p2id[I[q, 1]])  # the positive is among the neighbours of some queries
    args = argparse.Namespace(negative_sample=4)
    negs = ddrv.GenerateNegativePassaageID(args, passages, q_answers, q2id, p2id, I, pos)
    hits = ddrv.validate(passages, q_answers, I, q2id, p2id)
    with open(os.path.join(GOLD, "dpr_postprocess.json"), "w") as f:
        json.dump({"texts": texts, "answers": answers, "has_answer": table, "seed": 4, "n_p": n_p, "n_q": n_q, "k": k,
                   "negatives": {str(k_): [int(x) for x in v] for k_, v in negs.items()}, "top_k_hits": hits,
                   "negative_sample": 4}, f)
    print("dpr golden: hit@k", hits[0], hits[-1], "has_answer trues", sum(map(sum, table)))


def golden_msmarco_mrr():
    """MS MARCO MRR@10 from the reference's own utils/msmarco_eval.py:109-139 (pure Python, runs unmodified) on seeded
    random rankings: pins ance_b200.evaluation.msmarco_mrr (notebook cell 8's `ms_mrr`)."""
    from utils.msmarco_eval import compute_metrics
    rng = np.random.default_rng(17)
    cases = []
    for n_q, n_judged, n_p in ((40, 50, 300), (7, 5, 30)):
        relevant = {int(q): [int(x) for x in rng.choice(n_p, size=int(rng.integers(1, 4)), replace=False)]
                    for q in rng.choice(1000, size=n_judged, replace=False)}
        ranked = {}
        qids = list(relevant)[:n_q // 2] + [int(q) for q in rng.integers(1000, 2000, size=n_q - n_q // 2)]
        for q in qids:
            lst = [int(x) for x in rng.permutation(n_p)[:int(rng.integers(3, 40))]]
            if q in relevant and rng.random() < 0.7:   # plant a relevant passage somewhere in the first 15 ranks
                lst[int(rng.integers(0, min(15, len(lst))))] = relevant[q][0]
            ranked[q] = (lst + [0] * 1000)[:1000]
        cases.append({"relevant": {str(k): v for k, v in relevant.items()}, "ranked": {str(k): v[:50] for k, v in ranked.items()},
                      "mrr10": compute_metrics(relevant, ranked)["MRR @10"]})
    with open(os.path.join(GOLD, "msmarco_mrr.json"), "w") as f:
        json.dump(cases, f)
    print("msmarco mrr golden:", [c["mrr10"] for c in cases])


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    stub_third_party()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    golden_search()
    golden_io()
    golden_postprocess()
    golden_dpr()
    golden_trainer_records()
    golden_msmarco_mrr()
    if "--no-encoders" not in sys.argv:
        golden_encoders()
        golden_trainer_losses()
    print("golden fixtures written to", GOLD)
