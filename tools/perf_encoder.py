"""Device-time breakdown of the encoder by kernel class (ance_profile_*), per batch shape."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from ance_b200 import _lib  # noqa: E402
from ance_b200.models import RobertaDot_NLL_LN  # noqa: E402
from oracle.encoder_oracle import random_roberta_state_dict  # noqa: E402
from tools.bringup_encoder import roberta_cfg  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    model = RobertaDot_NLL_LN(roberta_cfg())
    model.load_state_dict(random_roberta_state_dict(seed=0), strict=True)
    model = model.to(dev).eval()
    out = open(ROOT / "gpurun_out" / "perf_encoder.jsonl", "a")
    import os
    if os.environ.get("ANCE_LN_ROWS"):
        model._encoder(dev).set_param("ln_rows_per_warp", int(os.environ["ANCE_LN_ROWS"]))
    shapes = [(512, 128), (128, 512), (1024, 64), (256, 256)]
    if len(sys.argv) > 1:
        shapes = [s for s in sys.argv[1].split(",")]
    for shp in shapes:
        ragged = isinstance(shp, str) and shp.endswith("r")   # "592x128r": lengths uniform in [L/8, L]
        B, L = (tuple(map(int, shp.rstrip("r").split("x"))) if isinstance(shp, str) else shp)
        g = torch.Generator(device=dev).manual_seed(0)
        ids = torch.randint(3, 50265, (B, L), device=dev, generator=g, dtype=torch.int32)
        lens = torch.full((B,), L, device=dev, dtype=torch.int32)
        if ragged:
            lens = torch.randint(max(1, L // 8), L + 1, (B,), device=dev, generator=g, dtype=torch.int32)
            ids = torch.where(torch.arange(L, device=dev)[None, :] < lens[:, None], ids, torch.ones_like(ids))
        for _ in range(3):
            model.encode_lens(ids, lens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 5
        e0.record()
        for _ in range(it):
            model.encode_lens(ids, lens)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / it
        _lib.profile_enable(True)
        _lib.profile_read(reset=True)
        for _ in range(it):
            model.encode_lens(ids, lens)
        prof = _lib.profile_read(reset=True)
        _lib.profile_enable(False)
        gemm_flop = B * L * 12 * 24 * 768 * 768 + B * 2 * 768 * 768
        attn_flop = B * 12 * 4 * 768 * L * L
        rec = {"B": B, "L": L, "ragged": ragged, "ms": ms, "seq_per_s": B / ms * 1e3, "tflops_total": (gemm_flop + attn_flop) / ms / 1e9,
               "gemm_ms": prof["encoder_gemm"][0] / it, "attn_ms": prof["attention"][0] / it,
               "norm_ms": prof["norm_embed"][0] / it,
               "gemm_tflops": gemm_flop / (prof["encoder_gemm"][0] / it) / 1e9,
               "attn_tflops": attn_flop / (prof["attention"][0] / it) / 1e9,
               "by_gemm_ms": {k: round(prof[k][0] / it, 4) for k in ("gemm_qkv", "gemm_out", "gemm_ffn1", "gemm_ffn2", "gemm_head")},
               "launches": {k: v[1] // it for k, v in prof.items() if v[1]}}
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
