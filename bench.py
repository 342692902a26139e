#!/usr/bin/env python
"""bench.py — ANN-refresh throughput (BASELINE.json metric) on N GPUs of one node.

One STEP = one slice of a full refresh at the refresh's own passage:query mix (8,841,823 : 502,939 = 17.6:1; the step uses 16:1):
    encode PB passages (rdot_nll, RoBERTa-base, L=128, full-length synthetic token ids)
  + encode QB train queries (L=64)
  + top-200 inner-product search of those QB queries against the resident 8,841,823 x 768 index
    (synthetic LayerNorm-like clustered rows; sharded i % N across ranks when N > 1,
    queries all-gathered, per-shard top-200 merged on rank 0).
value = (passages encoded + queries searched) per second, whole job.  `stages` breaks it down into
the two rates the metric names (passages encoded/s, queries top-200/s).

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PASSAGES = 8841823
N_QUERIES = 502939
DIM = 768
L_P, L_Q, TOPK = 128, 64, 200
METRIC = "ANN-refresh throughput: passages encoded/sec + queries top-200/sec, 8.8M x 768"
UNIT = "passages+queries/s"
FLOP_SEQ = lambda L: 12 * (24 * 768 * 768 * L + 4 * 768 * L * L) + 2 * 768 * 768
GEMM_FLOP_SEQ = lambda L: 12 * 24 * 768 * 768 * L + 2 * 768 * 768               # the GEMM kernel's share
# Last-layer pruning: out-proj + FFN of the last layer run on the CLS row only (result-identical), so
# 18 * 768^2 * (L - 1) FLOP per sequence are NOT executed.  Fractions of peak are computed from EXECUTED flops.
PRUNED_FLOP_SEQ = lambda L: 18 * 768 * 768 * (L - 1)


def roberta_cfg():
    from transformers import RobertaConfig
    return RobertaConfig(vocab_size=50265, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                         intermediate_size=3072, max_position_embeddings=514, type_vocab_size=1, layer_norm_eps=1e-5,
                         pad_token_id=1, bos_token_id=0, eos_token_id=2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"bf16_tflops": j.get("bf16_tflops_sustained", j.get("bf16_tflops")), "hbm_gbs": j.get("hbm_gbs"),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md, sustained)"}


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


_ORC = None


# =============================================================================================
# reference arm / CPU baseline: the reference's arithmetic (oracle port) on the host cores
# =============================================================================================
_SEARCH_SAMPLE = {}


def cpu_step_sample(threads, n_p=32, n_q=16, search_q=64, search_rows=262144):
    """Time a bounded sample of one step on the CPU and extrapolate to the step's unit counts.
    Returns (passages/s, query-encodes/s, search queries/s at N = 8,841,823, seconds spent)."""
    from oracle import flat_ip_oracle
    from oracle.encoder_oracle import RobertaDotOracle, random_roberta_state_dict
    torch.set_num_threads(threads)
    global _ORC
    if _ORC is None:
        _ORC = RobertaDotOracle(random_roberta_state_dict(seed=0))  # weights: setup, not timed
    orc = _ORC
    t_all = time.time()
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, 50265, (n_p, L_P), generator=g)
    orc.body_emb(ids[:2], torch.ones_like(ids[:2]))  # warm-up
    t0 = time.time()
    for s in range(0, n_p, 16):  # the reference's per_gpu_eval_batch_size in the shipped scripts
        orc.body_emb(ids[s:s + 16], torch.ones_like(ids[s:s + 16]))
    rate_p = n_p / (time.time() - t0)
    qids = torch.randint(3, 50265, (n_q, L_Q), generator=g)
    t0 = time.time()
    orc.query_emb(qids, torch.ones_like(qids))
    rate_q = n_q / (time.time() - t0)
    key = (search_rows, search_q)
    if key not in _SEARCH_SAMPLE:   # synthetic operands: setup, generated once per process, not timed
        rng = np.random.default_rng(0)
        _SEARCH_SAMPLE.clear()
        _SEARCH_SAMPLE[key] = (rng.standard_normal((search_rows, DIM), dtype=np.float32),
                               rng.standard_normal((search_q, DIM), dtype=np.float32))
    P, Q = _SEARCH_SAMPLE[key]
    t0 = time.time()
    flat_ip_oracle.search(P, Q, TOPK, slack=64, q_block=search_q, p_block=65536)
    qps_slice = search_q / (time.time() - t0)
    qps_full = qps_slice * search_rows / N_PASSAGES
    return rate_p, rate_q, qps_full, time.time() - t_all


# bounded samples of one step for the CPU arm: ~15 s of host work inside the default bench run, ~8 s per step of
# `--impl reference` (K + W steps must end within a few minutes)
CPU_BASELINE_SAMPLE = dict(n_p=192, n_q=48, search_q=128, search_rows=524288)
REF_STEP_SAMPLE = dict(n_p=96, n_q=24, search_q=64, search_rows=524288)
if os.environ.get("ANCE_BENCH_TINY_CPU"):   # contract tests only (tests/test_bench_contract.py)
    CPU_BASELINE_SAMPLE = REF_STEP_SAMPLE = dict(n_p=2, n_q=2, search_q=4, search_rows=4096)


def sample_text(sm):
    return ("%d passages L=%d + %d queries L=%d through the oracle port of the reference's HF-RoBERTa eager fp32 path (batch 16); "
            "%d queries x %s rows blocked fp32 sgemm + top-%d (faiss IndexFlatIP arithmetic), scaled linearly to N=%s"
            % (sm["n_p"], L_P, sm["n_q"], L_Q, sm["search_q"], format(sm["search_rows"], ","), TOPK, format(N_PASSAGES, ",")))


def cpu_threads():
    """The reference pins its CPU search to 16 OpenMP threads (run_ann_data_gen.py:269); small-batch fp32
    matmuls stop scaling (and regress) far below the 128+ hardware threads of a B200 host."""
    return min(os.cpu_count() or 1, 16)


def cpu_value(pb, qb, rate_p, rate_q, qps_full):
    t = pb / rate_p + qb / rate_q + qb / qps_full
    return (pb + qb) / t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = cpu_threads()
    pb, qb = args.passages_per_step, args.queries_per_step
    vals, spent = [], 0.0
    for i in range(args.warmup + args.steps):
        rate_p, rate_q, qps, dt = cpu_step_sample(threads, **REF_STEP_SAMPLE)
        spent += dt
        if i >= args.warmup:
            vals.append(cpu_value(pb, qb, rate_p, rate_q, qps))
    v = float(np.mean(vals))
    sample = "per step: " + sample_text(REF_STEP_SAMPLE) + "; extrapolated to the step's %d passages + %d queries; %.0f s of CPU work per step" % (
        pb, qb, spent / max(1, args.warmup + args.steps))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (pb + qb) / v * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: MS MARCO passage 8.8M, rdot_nll seq_len=128, encode + top-200",
            "index_rows": N_PASSAGES, "dim": DIM, "topk": TOPK, "passages_per_step_per_gpu": args.passages_per_step,
            "queries_per_step_per_gpu": args.queries_per_step, "passage_len": L_P, "query_len": L_Q,
            "parallelism": "rows i%%%d per rank, all-gather queries, host merge" % world if world > 1 else "single GPU",
            "search_operand": args.search_operand,
            "l2": "inputs larger than L2 (index 13.6 GB 16-bit + 27 GB fp32; ~0.7 GB of activations per encoder pass)"}


# =============================================================================================
# B200 arm
# =============================================================================================
def synth_index_rows(n, dev, seed, cent):
    g = torch.Generator(device=dev).manual_seed(seed)
    for s in range(0, n, 1 << 20):
        e = min(n, s + (1 << 20))
        x = 0.5 * torch.randn(e - s, DIM, device=dev, generator=g) + \
            0.5 * cent[torch.randint(0, cent.shape[0], (e - s,), device=dev, generator=g)]
        yield (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True, unbiased=False)


def run_b200(args):
    import torch.distributed as dist
    from ance_b200 import _lib
    from ance_b200.models import RobertaDot_NLL_LN
    from ance_b200.search import IndexFlatIP, merge_topk_host
    from oracle.encoder_oracle import random_roberta_state_dict  # seeded synthetic weights only (no oracle compute)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl")
    pb, qb = args.passages_per_step, args.queries_per_step

    model = RobertaDot_NLL_LN(roberta_cfg())
    model.load_state_dict(random_roberta_state_dict(seed=0), strict=True)
    model = model.to(dev).eval()
    # this rank's shard of the synthetic corpus
    n_local = len(range(rank, N_PASSAGES, world))
    offset = sum(len(range(r, N_PASSAGES, world)) for r in range(rank))
    index = IndexFlatIP(DIM, capacity=n_local, device=dev, operand=args.search_operand)
    cent = torch.randn(1024, DIM, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    for x in synth_index_rows(n_local, dev, 1234 + rank, cent):
        index.add(x)
    del x
    torch.cuda.empty_cache()

    # synthetic token ids: HOST pinned (e2e) and device-resident copies (value)
    g = torch.Generator().manual_seed(100 + rank)
    p_ids_h = torch.randint(3, 50265, (pb, L_P), generator=g, dtype=torch.int32).pin_memory()
    q_ids_h = torch.randint(3, 50265, (qb, L_Q), generator=g, dtype=torch.int32).pin_memory()
    p_ids_h[:, 0], q_ids_h[:, 0] = 0, 0
    p_len_h = torch.full((pb,), L_P, dtype=torch.int32).pin_memory()
    q_len_h = torch.full((qb,), L_Q, dtype=torch.int32).pin_memory()
    p_ids_d, q_ids_d, p_len_d, q_len_d = (t.to(dev) for t in (p_ids_h, q_ids_h, p_len_h, q_len_h))
    D_h = torch.empty((qb * world, TOPK), dtype=torch.float32).pin_memory()
    I_h = torch.empty((qb * world, TOPK), dtype=torch.int64).pin_memory()

    def step(host: bool):
        if host:
            pi, pl = p_ids_h.to(dev, non_blocking=True), p_len_h.to(dev, non_blocking=True)
            qi, ql = q_ids_h.to(dev, non_blocking=True), q_len_h.to(dev, non_blocking=True)
        else:
            pi, pl, qi, ql = p_ids_d, p_len_d, q_ids_d, q_len_d
        model.encode_lens(pi, pl)               # passages of this slice (rows stay in HBM)
        q = model.encode_lens(qi, ql)
        if world > 1:
            q_all = torch.empty((qb * world, DIM), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(q_all, q.contiguous())
        else:
            q_all = q
        D, I = index.search_device(q_all.contiguous(), TOPK, row_offset=offset)
        if world > 1:
            if rank == 0:
                Ds = [torch.empty_like(D) for _ in range(world)]
                Is = [torch.empty_like(I) for _ in range(world)]
                dist.gather(D, Ds, dst=0)
                dist.gather(I, Is, dst=0)
                # the k-way merge of the per-shard lists is part of the job: it is inside the timed region of
                # `value` as well as of `e2e`
                _, Im = merge_topk_host([d.cpu().numpy() for d in Ds], [i.cpu().numpy() for i in Is], TOPK)
                return Im
            dist.gather(D, None, dst=0)
            dist.gather(I, None, dst=0)
            return None
        if host:
            D_h.copy_(D, non_blocking=True)
            I_h.copy_(I, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return I_h
        return I

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(host: bool, steps: int):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            step(host)
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) if not host else (time.time() - t0) * 1e3  # e2e includes host work: wall clock
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    for _ in range(max(3, args.warmup)):
        step(False)
    sync()
    launches0 = _lib.load().ance_launch_count()
    _lib.profile_enable(True)
    _lib.profile_read(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(False, args.steps)
    clocks = sampler.stop() if sampler else None
    prof = _lib.profile_read(reset=True)
    _lib.profile_enable(False)
    launches = _lib.load().ance_launch_count() - launches0
    st = index.stats()
    step(True)
    ms_e2e = timed(True, max(2, args.steps // 2))
    # practical figure: MS-MARCO-like passage lengths ~ clipped N(76, 28), encoded with length
    # buckets (no FLOPs on all-padding tails).  Reported beside, never inside, `value`.
    gl = torch.Generator().manual_seed(5)
    mlens = torch.clamp(torch.normal(76.0, 28.0, (pb,), generator=gl).round(), 8, L_P).to(torch.int32).to(dev)
    for _ in range(2):
        model.encode_lens_bucketed(p_ids_d, mlens)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        model.encode_lens_bucketed(p_ids_d, mlens)
    e1.record()
    sync()
    ms_marco = e0.elapsed_time(e1) / 2

    if rank != 0:
        return
    units = (pb + qb) * world
    pk = peaks()
    gemm_ms, gemm_n = prof["encoder_gemm"]
    pruned = args.steps * (pb * PRUNED_FLOP_SEQ(L_P) + qb * PRUNED_FLOP_SEQ(L_Q))
    gemm_flop = args.steps * (pb * GEMM_FLOP_SEQ(L_P) + qb * GEMM_FLOP_SEQ(L_Q)) - pruned  # executed
    gemm_tf = gemm_flop / gemm_ms / 1e9
    coarse_ms, coarse_n = prof["coarse_search"]
    coarse_tf = args.steps * 2.0 * qb * world * n_local * DIM / coarse_ms / 1e9 if coarse_ms else None
    enc_ms = gemm_ms + prof["attention"][0] + prof["norm_embed"][0]
    srch_ms = prof["quantize"][0] + coarse_ms + prof["rescore"][0] + prof["exact"][0]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_gemm_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    out = {
        "metric": METRIC, "value": units / ms * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, world),
        "stages": {
            "passages_encoded_per_s": (pb + qb * L_Q / L_P) * args.steps / enc_ms * 1e3 * world,
            "queries_top200_per_s": qb * world * args.steps / srch_ms * 1e3,
            "encode_ms_per_step": enc_ms / args.steps, "search_ms_per_step": srch_ms / args.steps,
            "encode_frac_of_bf16_peak": ((args.steps * (pb * FLOP_SEQ(L_P) + qb * FLOP_SEQ(L_Q)) - pruned) / enc_ms / 1e9) / pk["bf16_tflops"],
            "encode_flop_per_passage": {"algorithmic": FLOP_SEQ(L_P), "executed": FLOP_SEQ(L_P) - PRUNED_FLOP_SEQ(L_P)},
            "search_coarse_tflops": coarse_tf,
            "search_coarse_frac_of_bf16_peak": coarse_tf / pk["bf16_tflops"] if coarse_tf else None,
            "search_stats": st,
            "passages_per_s_marco_like_lengths": pb / ms_marco * 1e3 * world,
        },
        "roofline": {"kernel": "tc05_gemm_kernel<EpStore> (encoder linear layers)", "bound": "tensor",
                     "achieved": gemm_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": gemm_tf / pk["bf16_tflops"],
                     "traffic": traffic, "peak_source": pk["source"] + " of measured", "launches": gemm_n,
                     "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                     "share_of_step": gemm_ms / (ms * args.steps)},
        "e2e": {"value": units / ms_e2e * 1e3, "unit": UNIT,
                "h2d_bytes_per_step": int(pb * (L_P * 4 + 4) + qb * (L_Q * 4 + 4)),
                "d2h_bytes_per_step": int(qb * world * TOPK * 12)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        rate_p, rate_q, qps, dt = cpu_step_sample(threads, **CPU_BASELINE_SAMPLE)
        out["cpu_baseline"] = {
            "value": cpu_value(pb, qb, rate_p, rate_q, qps), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": sample_text(CPU_BASELINE_SAMPLE) + "; %d threads (the reference pins faiss to 16, "
                      "run_ann_data_gen.py:269); %.0f s of CPU work" % (threads, dt),
            "passages_per_s": rate_p, "queries_top200_per_s": qps}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--passages_per_step", type=int, default=37888)   # 64 encoder passes of 592 x 128 tokens
    ap.add_argument("--queries_per_step", type=int, default=2368)     # 2 encoder passes of 1184 x 64 tokens (16:1)
    ap.add_argument("--search_operand", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
