"""Controlled A/B of the encoder's 16-bit storage format on one box: fp16 and bf16 encoders alternate (same process, same
inputs), several rounds each, so that box / thermal state does not decide the comparison.  -> gpurun_out/ab_operand.jsonl"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from ance_b200 import _lib  # noqa: E402
from ance_b200.models import RobertaDot_NLL_LN  # noqa: E402
from ance_b200.synthetic import random_roberta_state_dict, roberta_base_config  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, L = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "592x128").split("x"))
    rounds, iters = 4, 40
    models = {}
    sd = random_roberta_state_dict(seed=0)
    for op in ("fp16", "bf16"):
        m = RobertaDot_NLL_LN(roberta_base_config())
        m.load_state_dict(sd, strict=True)
        m.encoder_operand = op
        models[op] = m.to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(0)
    ids = torch.randint(3, 50265, (B, L), device=dev, generator=g, dtype=torch.int32)
    lens = torch.full((B,), L, device=dev, dtype=torch.int32)
    for m in models.values():
        for _ in range(5):
            m.encode_lens(ids, lens)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    res = {op: [] for op in models}
    for r in range(rounds):
        for op, m in models.items():
            _lib.profile_read(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                m.encode_lens(ids, lens)
            e1.record()
            torch.cuda.synchronize()
            prof = _lib.profile_read(reset=True)
            res[op].append({"ms": e0.elapsed_time(e1) / iters, "gemm": prof["encoder_gemm"][0] / iters,
                            "attn": prof["attention"][0] / iters, "norm": prof["norm_embed"][0] / iters})
    out = {"B": B, "L": L, "rounds": rounds, "iters": iters}
    for op, rs in res.items():
        out[op] = {k: sum(x[k] for x in rs) / len(rs) for k in rs[0]}
        out[op]["per_round_ms"] = [round(x["ms"], 4) for x in rs]
    out["fp16_over_bf16"] = {k: out["fp16"][k] / out["bf16"][k] for k in ("ms", "gemm", "attn", "norm")}
    print(json.dumps(out))
    with open(ROOT / "gpurun_out" / "ab_operand.jsonl", "a") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    main()
