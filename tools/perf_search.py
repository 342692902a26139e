"""Coarse-search throughput in the tensor regime (many queries) and the bench regime, with the device-time profile."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from ance_b200 import _lib  # noqa: E402
from ance_b200.search import IndexFlatIP  # noqa: E402
from tools.bringup_search import make_data  # noqa: E402


def main():
    N = int(sys.argv[1])
    nqs = [int(x) for x in sys.argv[2].split(",")]
    fmt = sys.argv[3] if len(sys.argv) > 3 else "bf16"
    dev = torch.device("cuda:0")
    P, Q = make_data(N, max(nqs), 768, "clustered", dev)
    idx = IndexFlatIP(768, capacity=N, operand=fmt)
    idx.add(P)
    del P
    torch.cuda.empty_cache()
    _lib.profile_enable(True)
    out = open(ROOT / "gpurun_out" / "perf_search.jsonl", "a")
    for cg in (2, 1):
        idx.set_param("cta_group", cg)
        for nq in nqs:
            q = Q[:nq].contiguous()
            for _ in range(2):
                idx.search_device(q, 200)
            torch.cuda.synchronize()
            _lib.profile_read(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx.search_device(q, 200)
            e1.record()
            torch.cuda.synchronize()
            prof = _lib.profile_read(reset=True)
            st = idx.stats()
            rec = {"N": N, "nq": nq, "fmt": fmt, "cta_group": cg, "ms": e0.elapsed_time(e1), "qps": nq / e0.elapsed_time(e1) * 1e3,
                   "coarse_ms": prof["coarse_search"][0], "coarse_tflops": 2.0 * nq * N * 768 / prof["coarse_search"][0] / 1e9,
                   "rescore_ms": prof["rescore"][0], "quant_ms": prof["quantize"][0], "exact_ms": prof["exact"][0], "stats": st}
            print(json.dumps(rec), flush=True)
            out.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
