"""N > 1 host logic on CPU: world_size-2 gloo process group, database row-sharded, one all_gather, merge.
The CUDA kernels are replaced by the oracle's CPU functions through ShardedIndexFlatL2's injection points;
what is exercised is the product's sharding / id-offset / gather-layout code."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_db, nq, d, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nano_vs_slam_b200.retrieval import ShardedIndexFlatL2, shard_bounds
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    db, q, planted = planted_retrieval_set(n_db, nq, d, k, seed=5)
    lo, hi = shard_bounds(n_db, world, rank)

    def local_search(shard, qq, kk, off):
        return glue_ref.flat_l2_search(shard, qq, kk, id_offset=off)

    def merge(Dp, Ip):
        return glue_ref.merge_shard_topk(list(Dp), list(Ip), Dp.shape[-1])

    idx = ShardedIndexFlatL2(d, n_db, device="cpu", local_search=local_search, merge=merge)
    assert (idx.lo, idx.hi) == (lo, hi) and idx.world == world and idx.rank == rank
    idx.add_local(db[lo:hi])
    D, I = idx.search(q, k)
    ok = torch.equal(I, planted) and bool((D[:, 1:] >= D[:, :-1]).all())
    Dr, Ir = glue_ref.flat_l2_search(db, q, k)
    ok = ok and torch.equal(I, Ir) and torch.allclose(D, Dr, rtol=1e-5, atol=1e-6)
    out[rank] = int(ok)
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("nq,k", [(24, 7), (23, 5)])  # nq*k even / odd: the packed (labels | distances) buffer pads
def test_sharded_index_world2_gloo(nq, k):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 1001, nq, 256, k, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_shard_bounds_cover_everything():
    from nano_vs_slam_b200.retrieval import shard_bounds

    for n, w in ((1_000_000, 8), (1001, 2), (7, 8), (125, 4)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
