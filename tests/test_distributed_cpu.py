"""world_size-2 (and 3) `gloo` runs of the sharded refresh logic on CPU: rank striding, the all-gather
of query rows in merged order, per-shard top-k with row offsets, all-to-all of the lists to the rank that owns
each query, host merge there, final gather of the merged labels.  The local
search is the CPU oracle (test infrastructure) — the point here is the host/collective logic."""
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


def _worker(rank, world, port, n_p, n_q, k, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ance_b200.data import EmbeddingCache, StreamingDataset, StridedBatchReader
        from ance_b200.drivers import run_ann_data_gen as drv
        rng = np.random.default_rng(11)
        P = rng.standard_normal((n_p, 24)).astype(np.float32)
        Q = rng.standard_normal((n_q, 24)).astype(np.float32)
        P[50:53] = P[3:6]  # ties that straddle shards
        mine_p = np.arange(rank, n_p, world)
        mine_q = np.arange(rank, n_q, world)
        p_loc, q_loc = torch.from_numpy(P[mine_p]), torch.from_numpy(Q[mine_q])
        # merged-order ids and rows
        p2id = drv.all_gather_ids(mine_p, torch.device("cpu"))
        q2id = drv.all_gather_ids(mine_q, torch.device("cpu"))
        assert p2id.tolist() == refresh_oracle.merged_embedding2id(n_p, world, 7)
        q_all = drv.all_gather_rows(q_loc)
        assert torch.equal(q_all, torch.from_numpy(Q[q2id]))

        def local_search(q, kk, row_offset):
            D, I = flat_ip_oracle.search_bruteforce(p_loc.numpy(), q.numpy(), kk)
            return torch.from_numpy(D), torch.from_numpy(np.where(I >= 0, I + row_offset, -1))

        I = drv.sharded_search(local_search, p_loc.shape[0], q_all, k, merge_threads=2)
        # several (ragged) query blocks in flight: the all-to-all / pipelined-merge path of a 503k-query refresh in small
        Ib = drv.sharded_search(local_search, p_loc.shape[0], q_all, k, merge_threads=1, query_block=5)
        own_I, own_q = drv.sharded_search(local_search, p_loc.shape[0], q_all, k, query_block=7, gather_to_rank0=False)
        _, Ig = flat_ip_oracle.search_bruteforce(P[p2id], Q[q2id], k)
        assert (own_I == Ig[own_q]).all()          # every rank holds the merged lists of the queries it owns
        counts = [None] * world
        dist.all_gather_object(counts, own_q.tolist())
        assert sorted(sum(counts, [])) == list(range(n_q))   # ... and every query is owned exactly once
        # two searches in flight (bench.py at N > 1: the merge of slice i runs while slice i + 1 is encoded and searched):
        # each owns its staging set; finishing in issue order gives the synchronous results
        pa = drv.sharded_search_start(local_search, p_loc.shape[0], q_all, k, query_block=5)
        pb = drv.sharded_search_start(local_search, p_loc.shape[0], q_all.flip(0).contiguous(), k, query_block=5)
        assert pa.sset is not pb.sset
        Ia, Ifl = pa.finish(), pb.finish()
        pc = drv.sharded_search_start(local_search, p_loc.shape[0], q_all, k, query_block=5)   # the sets are free again
        assert pc.sset in (pa.sset, pb.sset)
        Ic = pc.finish()
        # edges: fewer queries than ranks, and none at all
        I1 = drv.sharded_search(local_search, p_loc.shape[0], q_all[:1].contiguous(), k)
        I0 = drv.sharded_search(local_search, p_loc.shape[0], q_all[:0].contiguous(), k)
        if rank == 0:
            assert (I1 == Ig[:1]).all() and I0.shape == (0, k)
            assert (I == Ig).all() and (Ib == Ig).all()
            assert (Ia == Ig).all() and (Ic == Ig).all() and (Ifl == Ig[::-1]).all()
            np.save(os.path.join(tmpdir, "I.npy"), I)
        else:
            assert I is None and Ib is None and Ia is None and Ifl is None and Ic is None
        # the reference's StreamingDataset stride under an initialised process group
        lens = np.ones(n_q, dtype=np.int32)
        base = os.path.join(tmpdir, f"cache{rank}")
        refresh_oracle.write_cache(base, lens, np.arange(n_q * 4, dtype=np.int32).reshape(n_q, 4))
        with EmbeddingCache(base) as c:
            got = [int(r[1]) for r in StreamingDataset(c, lambda e, i: [(e, i)])]
        assert got == mine_q.tolist()
        seen = [x for _, _, idx in StridedBatchReader(EmbeddingCache(base), 5, rank, world, pin=False) for x in idx.tolist()]
        assert seen == mine_q.tolist()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_p,n_q,k", [(2, 203, 31, 10), (3, 100, 8, 40)])
def test_sharded_search_gloo(tmp_path, world, n_p, n_q, k):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_p, n_q, k, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "I.npy")
