"""GPU bring-up of the flat-IP search: exact path vs a torch fp64 reference, coarse+rescore path
vs the exact path (must be bit-identical), timing at growing sizes.  One subprocess per case."""
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def make_data(N, nq, d, kind, dev):
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    cent = torch.randn(1024, d, device=dev, generator=g)

    def gen(n, seed):
        out = torch.empty(n, d, device=dev)
        g.manual_seed(seed)
        for s in range(0, n, 1 << 20):
            e = min(n, s + (1 << 20))
            x = torch.randn(e - s, d, device=dev, generator=g)
            if kind == "clustered":
                a = torch.randint(0, 1024, (e - s,), device=dev, generator=g)
                x = 0.5 * x + 0.5 * cent[a]
            # LayerNorm-like rows (mean 0, var 1): what the head's nn.LayerNorm(768) emits (models.py:153)
            x = (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True, unbiased=False)
            out[s:e] = x
        return out

    P = gen(N, 1234)
    Q = gen(nq, 4321)
    if kind == "dups":  # planted duplicate rows -> exact ties
        P[N // 2:N // 2 + 4096] = P[:4096]
    return P, Q


def child(N, nq, k, fmt, cg, kind, check, kprime):
    import torch
    from ance_b200.search import IndexFlatIP
    dev = torch.device("cuda:0")
    d = 768
    P, Q = make_data(N, nq, d, kind, dev)
    idx = IndexFlatIP(d, capacity=N, operand="bf16" if fmt == 1 else "fp16")
    idx.add(P)
    if not check:
        del P
        torch.cuda.empty_cache()
    idx.set_param("cta_group", cg)
    if kprime:
        idx.set_param("kprime", kprime)
    res = {}
    if check:
        t0 = time.time()
        De, Ie = idx.search_device(Q, k, exact=True)
        torch.cuda.synchronize()
        res["exact_s"] = round(time.time() - t0, 3)
        # torch fp64 reference on a slice of queries
        nref = min(nq, 256)
        S = (Q[:nref].double() @ P.double().t()).float()
        Ss, Si = torch.sort(S, dim=1, descending=True, stable=True)
        res["exact_I_match"] = bool((Si[:, :k] == Ie[:nref]).all().item())
        res["exact_D_match"] = bool((Ss[:, :k] == De[:nref]).all().item())
        if not res["exact_I_match"]:
            bad = (Si[:, :k] != Ie[:nref]).sum().item()
            res["exact_I_mismatch_count"] = bad
            res["exact_D_maxdiff"] = (Ss[:, :k] - De[:nref]).abs().max().item()
    print("data+index ready", time.time(), file=sys.stderr, flush=True)
    D, I = idx.search_device(Q, k)
    torch.cuda.synchronize()
    print("first search done", time.time(), file=sys.stderr, flush=True)
    st = idx.stats()
    res["stats"] = st
    if check:
        res["coarse_I_match"] = bool((I == Ie).all().item())
        res["coarse_D_match"] = bool((D == De).all().item())
        if not res["coarse_I_match"]:
            res["coarse_I_mismatch_rows"] = int((I != Ie).any(1).sum().item())
    # timing (includes the exact fallback launches, which exit immediately when nothing is flagged)
    for _ in range(2):
        idx.search_device(Q, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 3
    e0.record()
    for _ in range(iters):
        idx.search_device(Q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res["ms"] = ms
    res["qps"] = nq / ms * 1e3
    res["tflops"] = 2.0 * nq * N * d / ms / 1e9
    print(json.dumps(res))


def main():
    OUT.mkdir(exist_ok=True)
    cases = [
        # N, nq, k, fmt, cg, kind, check, kprime
        (20000, 100, 20, 1, 1, "iid", 1, 0),
        (100000, 300, 100, 1, 1, "iid", 1, 0),
        (100000, 300, 200, 1, 1, "clustered", 1, 0),
        (100000, 300, 200, 0, 1, "clustered", 1, 0),
        (100000, 300, 200, 1, 2, "clustered", 1, 0),
        (100000, 300, 200, 1, 1, "dups", 1, 0),
        (300000, 1000, 200, 1, 1, "clustered", 1, 512),
        (1000000, 8192, 200, 1, 1, "clustered", 0, 0),
        (1000000, 8192, 200, 0, 1, "clustered", 0, 0),
        (1000000, 8192, 200, 1, 2, "clustered", 0, 0),
        (1000000, 37888, 200, 1, 1, "clustered", 0, 0),
        (1000000, 37888, 200, 1, 2, "clustered", 0, 0),
        (8841823, 37888, 200, 1, 1, "clustered", 0, 0),
        (8841823, 6980, 100, 1, 1, "clustered", 0, 0),
        (2000000, 37888, 200, 1, 1, "clustered", 0, 0),
        (4000000, 37888, 200, 1, 1, "clustered", 0, 0),
        (8841823, 18944, 200, 1, 1, "clustered", 0, 0),
        (8841823, 37888, 200, 1, 1, "iid", 0, 0),
    ]
    if len(sys.argv) > 1:
        cases = [cases[int(i)] for i in sys.argv[1].split(",")]
    with open(OUT / "bringup_search.jsonl", "a") as f:
        for c in cases:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "child", *map(str, c)], capture_output=True, text=True,
                                   timeout=240)
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
                err = e.stderr.decode("utf-8", "replace") if isinstance(e.stderr, bytes) else str(e.stderr)
                res = {"ok": False, "timeout": True, "stderr": err[-2000:]}
            res["case"] = c
            res["wall_s"] = round(time.time() - t0, 1)
            print(json.dumps(res), flush=True)
            f.write(json.dumps(res) + "\n")
            f.flush()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        a = sys.argv[2:]
        child(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), a[5], int(a[6]), int(a[7]))
    else:
        main()
