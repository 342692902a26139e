"""A/B of encoder build switches on ONE box (box-to-box variance is ~5 %, so variants must share a box):

    python tools/ab_encoder.py --shape 592x128 --reps 3 "" "ANCE_B200_EPI_MASK=15" "ANCE_B200_GELU=1"

Every variant is a set of NAME=VALUE environment assignments (comma separated, "" = defaults); variants are run
interleaved, each in its own process through tools/perf_encoder.py, and the median per kernel class is printed."""
import argparse
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="592x128")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    res = {v: [] for v in a.variants}
    for _ in range(a.reps):
        for v in a.variants:
            env = dict(os.environ)
            for kv in filter(None, v.split(",")):
                k, _, val = kv.partition("=")
                env[k] = val
            r = subprocess.run([sys.executable, str(ROOT / "tools" / "perf_encoder.py"), a.shape], env=env,
                               capture_output=True, text=True, timeout=300)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode != 0 or not line:
                print(f"variant {v!r} failed: {r.stderr[-400:]}", flush=True)
                continue
            res[v].append(json.loads(line[-1]))
    keys = ["ms", "gemm_ms", "attn_ms", "norm_ms"]
    print(f"{'variant':40s} " + " ".join(f"{k:>9s}" for k in keys) + "   qkv    out   ffn1   ffn2")
    for v, rs in res.items():
        if not rs:
            continue
        med = lambda f: statistics.median(f(r) for r in rs)   # noqa: E731
        g = [med(lambda r, c=c: r["by_gemm_ms"][c]) for c in ("gemm_qkv", "gemm_out", "gemm_ffn1", "gemm_ffn2")]
        print(f"{(v or '(defaults)'):40s} " + " ".join(f"{med(lambda r, k=k: r[k]):9.3f}" for k in keys) + "  " +
              " ".join(f"{x:6.3f}" for x in g), flush=True)


if __name__ == "__main__":
    main()
