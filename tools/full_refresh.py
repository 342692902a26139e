#!/usr/bin/env python
"""One full ANN refresh through the drop-in driver at BASELINE.json's sizes (north_star "Target"):

    synthetic MS-MARCO-shaped data_dir (8,841,823 passages L=128, 502,939 train + 6,980 dev queries L=64) and a
    seeded-random RoBERTa-base checkpoint  ->  ance_b200.drivers.run_ann_data_gen (the reference's CLI:
    --topk_training 200 --ann_chunk_factor 1 --negative_sample 20)  ->  ann_training_data_0 + ann_ndcg_0

    python tools/full_refresh.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
           tools/full_refresh.py                                   # 8 GPUs, rows i % 8

Writes gpurun_out/full_refresh_n{W}.json: wall-clock per stage (encode / search / post-processing) as the driver
measures them, and each stage's fraction of the measured sustained bf16 peak (MEASURED_PEAKS.json) computed from
EXECUTED flops (last-layer pruning and length buckets are not claimed).  The data generation and the model load are
reported separately; they are not part of a refresh's steady state (the caches exist once, the checkpoint is read per
refresh).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def flop_seq(L):
    return 12 * (24 * 768 * 768 * L + 4 * 768 * L * L) + 2 * 768 * 768


def pruned(L):
    return 18 * 768 * 768 * (L - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_passages", type=int, default=8841823)
    ap.add_argument("--n_queries", type=int, default=502939)
    ap.add_argument("--n_dev", type=int, default=6980)
    ap.add_argument("--lengths", default="full", choices=["full", "marco"],
                    help="full: every passage has 128 real tokens (the roofline regime, SURVEY.md 8d); marco: clipped N(76, 28)")
    ap.add_argument("--work_dir", default="", help="default: /dev/shm/ance_full_refresh when it has room, else /tmp/ance_full_refresh")
    ap.add_argument("--topk_training", type=int, default=200)
    ap.add_argument("--negative_sample", type=int, default=20)
    ap.add_argument("--n_layer", type=int, default=12)
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()

    from ance_b200 import synthetic
    from ance_b200.drivers import run_ann_data_gen as drv

    W = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if W > 1:
        dist.init_process_group("nccl")
    barrier = (lambda: dist.barrier()) if W > 1 else (lambda: None)

    if not a.work_dir:
        need = a.n_passages * 516 + (a.n_queries + a.n_dev) * 260 + (2 << 30)
        try:
            shm_ok = shutil.disk_usage("/dev/shm").free > need
        except OSError:
            shm_ok = False
        a.work_dir = "/dev/shm/ance_full_refresh" if shm_ok else "/tmp/ance_full_refresh"
    data, ckpt, out = (os.path.join(a.work_dir, x) for x in ("data", "init_model", "ann"))
    t0 = time.time()
    if rank == 0:
        shutil.rmtree(a.work_dir, ignore_errors=True)
        os.makedirs(data)
    barrier()
    synthetic.write_marco_like_dir(data, a.n_passages, a.n_queries, a.n_dev, full_length_passages=(a.lengths == "full"),
                                   part=rank, n_parts=W, barrier=barrier)
    if rank == 0:
        synthetic.write_checkpoint(ckpt, seed=0, n_layer=a.n_layer)
    barrier()
    t_data = time.time() - t0

    argv = ["--data_dir", data, "--training_dir", os.path.join(a.work_dir, "no_training_dir"), "--init_model_dir", ckpt,
            "--model_type", "rdot_nll", "--output_dir", out, "--cache_dir", os.path.join(a.work_dir, "cache"),
            "--end_output_num", "0", "--max_seq_length", "128", "--max_query_length", "64",
            "--per_gpu_eval_batch_size", "16", "--topk_training", str(a.topk_training), "--negative_sample",
            str(a.negative_sample), "--ann_chunk_factor", "1", "--seed", "0"]
    if a.lengths == "full":
        argv.append("--no_length_buckets")   # nothing to bucket: every passage is 128 tokens (queries keep their padding too)
    args = drv.get_arguments(argv)
    if W > 1:
        args.local_rank = local
    drv.set_env(args)
    t1 = time.time()
    _, _, model = drv.load_model(args, ckpt)
    backend = drv.B200Backend(args, model)
    torch.cuda.synchronize()
    t_load = time.time() - t1
    barrier()
    t2 = time.time()
    drv.ann_data_gen(args, backend=backend)
    torch.cuda.synchronize()
    barrier()
    t_total = time.time() - t2
    if rank != 0:
        return
    tm = args.last_refresh_timing
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    # executed FLOPs of the encode stage (dense, padded to the cache length unless bucketed; last layer pruned)
    if a.lengths == "full":
        enc_flop = a.n_passages * (flop_seq(128) - pruned(128)) + (a.n_queries + a.n_dev) * (flop_seq(64) - pruned(64))
        enc_note = "dense L=128 passages, L=64 queries (no buckets), last-layer pruning subtracted"
    else:
        enc_flop = None
        enc_note = "length buckets: executed FLOPs depend on the bucket histogram; passages/s is the figure to read"
    srch_flop = 2.0 * (a.n_queries + a.n_dev) * a.n_passages * 768
    lines = sum(1 for _ in open(os.path.join(out, "ann_training_data_0")))
    res = {
        "what": "one full refresh through ance_b200.drivers.run_ann_data_gen", "n_gpus": W, "lengths": a.lengths,
        "n_passages": a.n_passages, "n_train_queries": a.n_queries, "n_dev_queries": a.n_dev,
        "topk_training": a.topk_training, "negative_sample": a.negative_sample, "n_layer": a.n_layer,
        "refresh_wall_s": t_total, "encode_s": tm["encode_s"], "search_s": tm["search_s"], "post_s": tm["post_s"],
        "stage_detail": tm.get("detail"),
        "passages_per_s": a.n_passages / tm["encode_s"] if tm["encode_s"] else None,   # (queries are inside encode_s too)
        "queries_per_s": (a.n_queries + a.n_dev) / tm["search_s"] if tm["search_s"] else None,
        "encode_frac_of_sustained_bf16_peak": (enc_flop / tm["encode_s"] / 1e12 / (peak * W)) if enc_flop else None,
        "search_frac_of_sustained_bf16_peak": srch_flop / tm["search_s"] / 1e12 / (peak * W),
        "peak_tflops_per_gpu": peak, "encode_flop_note": enc_note,
        "search_stats_last_call": tm.get("search_stats"),
        "setup": {"data_generation_s": t_data, "model_load_s": t_load},
        "ann_training_data_lines": lines, "ndcg": json.load(open(os.path.join(out, "ann_ndcg_0"))),
        "data": "synthetic (ance_b200/synthetic.py), seeded-random RoBERTa-base",
    }
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = "full_refresh_n%d%s.json" % (W, ("_" + a.tag) if a.tag else "")
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)
    print(json.dumps(res))
    if not a.keep:
        shutil.rmtree(a.work_dir, ignore_errors=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
