"""GPU bring-up of the encoder: per-layer hidden states vs the CPU oracle, final embeddings vs the
golden outputs of the reference classes, throughput.  One subprocess per case."""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def roberta_cfg():
    from transformers import RobertaConfig
    return RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                         intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                         pad_token_id=1, bos_token_id=0, eos_token_id=2)


def stats(a, b):
    import torch
    a, b = a.float().cpu(), b.float().cpu()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
    return {"max_abs": (a - b).abs().max().item(), "min_cos": cos}


def child(case):
    import numpy as np
    import torch
    from ance_b200.models import BiEncoder, RobertaDot_CLF_ANN_NLL_MultiChunk, RobertaDot_NLL_LN
    from oracle.encoder_oracle import BiEncoderOracle, RobertaDotOracle, random_roberta_state_dict
    dev = torch.device("cuda:0")
    res = {}
    if case in ("rdot", "rdot_q", "multi"):
        sd = random_roberta_state_dict(seed=0)
        cls = RobertaDot_CLF_ANN_NLL_MultiChunk if case == "multi" else RobertaDot_NLL_LN
        model = cls(roberta_cfg())
        model.load_state_dict(sd, strict=True)
        model = model.to(dev).eval()
        orc = RobertaDotOracle(sd)
        if case == "multi":
            g = np.load(ROOT / "tests/golden/encoder_multi_chunk.npz")
            ids, lens, gold = g["ids"], g["lens"], g["emb"]
            mask = (np.arange(ids.shape[1])[None, :] < lens[:, None])
            enc = model._encoder(dev)
            enc.enable_debug()
            emb = model.body_emb(torch.from_numpy(ids).to(dev).long(), torch.from_numpy(mask).to(dev).long())
            torch.cuda.synchronize()
            real = g["real_chunk"]
            res["vs_reference_real_chunks"] = stats(emb[torch.from_numpy(real)], torch.from_numpy(gold[real]))
            res["allpad_vs_oracle_2_3_0"] = stats(emb[1, 2:4], torch.from_numpy(g["allpad_oracle_2_3_0"])[None].expand(2, -1))
            res["allpad_identical"] = bool((emb[1, 2] == emb[1, 3]).all().item())
            emb2 = model.encode_lens_multi_chunk(torch.from_numpy(ids).to(dev), torch.from_numpy(lens).to(dev))
            res["lens_path_equal"] = bool((emb2 == emb).all().item())
            ids2 = ids.reshape(-1, 512)
            mask2 = mask.reshape(-1, 512)
        else:
            g = np.load(ROOT / "tests/golden/encoder_rdot_nll.npz")
            if case == "rdot":
                ids, lens, gold = g["ids"], g["lens"], g["emb"]
            else:
                ids, lens, gold = g["qids"], g["qlens"], g["qemb"]
            mask = (np.arange(ids.shape[1])[None, :] < lens[:, None])
            enc = model._encoder(dev)
            enc.enable_debug()
            emb = model.body_emb(torch.from_numpy(ids).to(dev).long(), torch.from_numpy(mask).to(dev).long())
            torch.cuda.synchronize()
            res["vs_reference"] = stats(emb, torch.from_numpy(gold))
            emb2 = model.encode_lens(torch.from_numpy(ids).to(dev), torch.from_numpy(lens).to(dev))
            res["lens_path_equal"] = bool((emb2 == emb).all().item())
            ids2, mask2 = ids, mask
        hs = orc.enc.hidden_states(torch.from_numpy(ids2), torch.from_numpy(mask2))
        n_tok = ids2.shape[0] * ids2.shape[1]
        per_layer = []
        m = torch.from_numpy(mask2.reshape(-1))
        for l in range(13):
            h = enc.hidden(l, n_tok).cpu()
            ref = hs[l].reshape(n_tok, -1)
            if l == 12:  # pruned last layer: CLS rows first
                h = h[:ids2.shape[0]]
                ref = hs[l][:, 0]
                m = torch.ones(ids2.shape[0], dtype=torch.bool)
            d = (h - ref).abs()
            per_layer.append({"layer": l, "max_abs_real_tokens": d[m].max().item(),
                              "max_abs_all": d.max().item(), "nan": int(torch.isnan(h).sum())})
        res["per_layer"] = per_layer
    elif case == "dpr":
        sdq = random_roberta_state_dict(seed=1, vocab=30522, max_pos=512, head=False, prefix="question_model.")
        sdc = random_roberta_state_dict(seed=2, vocab=30522, max_pos=512, head=False, prefix="ctx_model.")
        model = BiEncoder()
        model.load_state_dict({**sdq, **sdc})
        model = model.to(dev).eval()
        g = np.load(ROOT / "tests/golden/encoder_dpr.npz")
        ids = torch.from_numpy(g["ids"]).to(dev)
        res["body_vs_reference"] = stats(model.body_emb(ids.long(), (ids != 0).long()), torch.from_numpy(g["body_emb"]))
        res["query_vs_reference"] = stats(model.query_emb(ids.long(), (ids != 0).long()), torch.from_numpy(g["query_emb"]))
    elif case.startswith("perf"):
        _, B, L = case.split(":")
        B, L = int(B), int(L)
        sd = random_roberta_state_dict(seed=0)
        model = RobertaDot_NLL_LN(roberta_cfg())
        model.load_state_dict(sd, strict=True)
        model = model.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(0)
        ids = torch.randint(3, 50265, (B, L), device=dev, generator=g, dtype=torch.int32)
        ids[:, 0] = 0
        lens = torch.full((B,), L, device=dev, dtype=torch.int32)
        for _ in range(3):
            model.encode_lens(ids, lens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5
        e0.record()
        for _ in range(it):
            out = model.encode_lens(ids, lens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        flop = B * (12 * (24 * 768 * 768 * L + 4 * 768 * L * L) + 2 * 768 * 768)
        res = {"ms": ms, "seq_per_s": B / ms * 1e3, "tflops": flop / ms / 1e9, "finite": bool(torch.isfinite(out).all())}
    print(json.dumps(res))


def main():
    OUT.mkdir(exist_ok=True)
    cases = ["rdot", "rdot_q", "multi", "dpr", "perf:512:128", "perf:128:512", "perf:1024:64"]
    if len(sys.argv) > 1:
        cases = sys.argv[1].split(",")
    with open(OUT / "bringup_encoder.jsonl", "a") as f:
        for c in cases:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "child", c], capture_output=True, text=True, timeout=300)
                line = (r.stdout.strip().splitlines() or ["{}"])[-1]
                try:
                    res = json.loads(line)
                except Exception:
                    res = {"ok": False, "stdout": r.stdout[-2000:]}
                res["rc"] = r.returncode
                if r.returncode != 0:
                    res["stderr"] = r.stderr[-2500:]
                    res["stdout"] = r.stdout[-1500:]
            except subprocess.TimeoutExpired as e:
                res = {"ok": False, "timeout": True, "stderr": (e.stderr or b"")[-2000:].decode("utf-8", "replace") if isinstance(e.stderr, bytes) else str(e.stderr)[-2000:]}
            res["case"] = c
            res["wall_s"] = round(time.time() - t0, 1)
            print(json.dumps(res), flush=True)
            f.write(json.dumps(res) + "\n")
            f.flush()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
    else:
        main()
