"""The refresher's whole host flow (`generate_new_ann`: stage order, rank striding, all-gathers in merged order, query
chunking, block-wise sharded search with a ragged last block, all-to-all + per-rank merge, array-form negatives, native line
writer, staged file names) on CPU under `gloo` with world sizes 1, 2 and 3.  The device work is replaced by a backend whose
"encoder" is a fixed random projection and whose "index" is the CPU search oracle (test infrastructure): what is checked is
that every world size writes the SAME ann_training_data / ann_ndcg, equal to the single-process oracle pipeline."""
import argparse
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import flat_ip_oracle, refresh_oracle


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeIndex:
    def __init__(self, rows):
        self.rows, self.ntotal = rows, rows.shape[0]

    def stats(self):
        return {"nq": 0}


class _FakeBackend:
    """encode(): embedding of record i = row i of a seeded table (so it does not depend on the rank that encodes it)."""

    def __init__(self, args, tables):
        self.args, self.tables = args, tables

    def encode(self, cache_path, is_query, build_index=False):
        from ance_b200.drivers.run_ann_data_gen import _world
        W, rank = _world()
        table = self.tables[os.path.basename(cache_path)]
        ids = np.arange(rank, table.shape[0], W, dtype=np.int64)
        rows = torch.from_numpy(table[ids])
        return (_FakeIndex(rows), rows, ids) if build_index else (rows, ids)

    def make_local_search(self, index):
        self.index = index

        def search(q, k, row_offset):
            D, I = flat_ip_oracle.search_bruteforce(index.rows.numpy(), q.numpy(), k)
            return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + row_offset, -1))
        return search


def _tables(n_p, n_q, n_dev):
    rng = np.random.default_rng(77)
    return {"passages": rng.standard_normal((n_p, 16)).astype(np.float32),
            "train-query": rng.standard_normal((n_q, 16)).astype(np.float32),
            "dev-query": rng.standard_normal((n_dev, 16)).astype(np.float32)}


def _worker(rank, world, port, out_dir, n_p, n_q, n_dev, chunk_factor, output_num):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ance_b200.drivers import run_ann_data_gen as drv
        drv.QUERY_BLOCK = 7                                     # several blocks, ragged last one
        rng = np.random.default_rng(5)
        train_pos = {q: int(rng.integers(0, n_p)) for q in range(n_q)}
        dev_pos = {q: {int(rng.integers(0, n_p)): 1} for q in range(n_dev)}
        args = argparse.Namespace(data_dir="unused", output_dir=out_dir, inference=False, ann_chunk_factor=chunk_factor,
                                  topk_training=12, negative_sample=4, ann_measure_topk_mrr=True, reference_sampling=False,
                                  seed=3, rank=rank, device=torch.device("cpu"))
        res = drv.generate_new_ann(args, output_num, "ckpt-x", train_pos, dev_pos, 0,
                                   backend=_FakeBackend(args, _tables(n_p, n_q, n_dev)))
        assert (res is None) == (rank != 0)
        if world > 1:
            dist.barrier()
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize("chunk_factor,output_num", [(1, 0), (3, 1)])
def test_generate_new_ann_is_world_size_invariant(tmp_path, chunk_factor, output_num):
    n_p, n_q, n_dev = 211, 45, 9
    files = {}
    for world in (1, 2, 3):
        out = tmp_path / f"w{world}"
        if world == 1:
            _worker(0, 1, 0, str(out), n_p, n_q, n_dev, chunk_factor, output_num)
        else:
            mp.spawn(_worker, args=(world, _free_port(), str(out), n_p, n_q, n_dev, chunk_factor, output_num), nprocs=world,
                     join=True)
        names = sorted(os.listdir(out))
        assert names == [f"ann_ndcg_{output_num}", f"ann_training_data_{output_num}"], names   # no staged leftovers
        files[world] = (sorted(open(out / f"ann_training_data_{output_num}").read().splitlines()),
                        json.load(open(out / f"ann_ndcg_{output_num}")))
    if chunk_factor == 1:
        assert files[1] == files[2] == files[3]
    # ... and equal to the oracle pipeline: exact search, first negative_sample + 1 neighbours.  With --ann_chunk_factor > 1 a
    # refresh covers a slice of the RANK-MAJOR merged query order (run_ann_data_gen.py:281-296), i.e. which queries belong
    # to it depends on the world size, as in the reference.
    t = _tables(n_p, n_q, n_dev)
    rng = np.random.default_rng(5)
    train_pos = {q: int(rng.integers(0, n_p)) for q in range(n_q)}
    for world in (1, 2, 3):
        merged = np.concatenate([np.arange(r, n_q, world) for r in range(world)])
        start, end = refresh_oracle.query_chunk(n_q, output_num, chunk_factor)
        qs = merged[start:end]
        _, I = flat_ip_oracle.search_bruteforce(t["passages"], t["train-query"][qs], 12)
        want = []
        for r, q in enumerate(qs.tolist()):
            negs = []
            for pid in I[r, :5]:
                if int(pid) != train_pos[q] and int(pid) not in negs and len(negs) < 4:
                    negs.append(int(pid))
            want.append("{}\t{}\t{}".format(q, train_pos[q], ",".join(map(str, negs))))
        assert files[world][0] == sorted(want), world
        assert files[world][1]["checkpoint"] == "ckpt-x" and 0.0 <= files[world][1]["ndcg"] <= 1.0
    assert files[1][1] == files[2][1] == files[3][1]          # the dev set is never chunked
