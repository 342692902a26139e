"""How often does the coarse pass fail to certify, by operand format / k' / corpus size?  (fallback off)"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from ance_b200 import _lib  # noqa: E402
from ance_b200.search import IndexFlatIP  # noqa: E402
from tools.bringup_search import make_data  # noqa: E402


def main():
    N = int(sys.argv[1])
    nq = int(sys.argv[2])
    kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["clustered"]
    dev = torch.device("cuda:0")
    out = open(ROOT / "gpurun_out" / "exp_certify.jsonl", "a")
    _lib.profile_enable(True)
    for kind in kinds:
        P, Q = make_data(N, nq, 768, kind, dev)
        for fmt in ("bf16", "fp16"):
            idx = IndexFlatIP(768, capacity=N, operand=fmt)
            idx.add(P)
            idx.set_param("exact_fallback", 0)
            idx.set_param("n_splits", 1)
            for k, kp in ((200, 288), (200, 512), (200, 992), (100, 160), (100, 256)):
                idx.set_param("kprime", kp)
                idx.search_device(Q, k)
                torch.cuda.synchronize()
                _lib.profile_read(reset=True)
                t0 = time.time()
                idx.search_device(Q, k)
                torch.cuda.synchronize()
                dt = time.time() - t0
                prof = _lib.profile_read(reset=True)
                st = idx.stats()
                rec = {"N": N, "nq": nq, "kind": kind, "fmt": fmt, "k": k, "kprime": kp,
                       "uncertified": st["n_uncertified"], "frac": st["n_uncertified"] / nq, "max_eps": st["max_eps"],
                       "wall_ms": dt * 1e3, "coarse_ms": prof["coarse_search"][0], "rescore_ms": prof["rescore"][0],
                       "quant_ms": prof["quantize"][0],
                       "coarse_tflops": 2.0 * nq * N * 768 / prof["coarse_search"][0] / 1e9}
                print(json.dumps(rec), flush=True)
                out.write(json.dumps(rec) + "\n")
                out.flush()
            del idx
            torch.cuda.empty_cache()
        del P, Q
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
