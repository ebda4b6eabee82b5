"""NetVLAD retrieval throughput (BASELINE config 5): Q queries vs an N-row, D-wide database, top-k.

The database is row-sharded over the ranks (N/world rows each, strong scaling: the global database is
fixed), queries are replicated, every rank searches its shard and ONE NCCL all_gather + merge produces the
global top-k.  Planted tie-free data (synthetic.planted_retrieval_set) is generated on the device; every
rank builds the same global set from the same seed and keeps only its rows, so results can be checked
against the planted answer without any CPU work.
"""
from __future__ import annotations

import os
import statistics

import torch


def shard_rows(world: int, rank: int, n_db: int = None) -> int:
    """Database rows of this rank's shard in run()."""
    from .retrieval import shard_bounds

    n_db = int(os.environ.get("NVS_RETR_NDB", 1_000_000)) if n_db is None else n_db
    lo, hi = shard_bounds(n_db, world, rank)
    return hi - lo


def run(dev, world: int, rank: int, n_db: int = None, n_q: int = None, dim: int = 4096, k: int = 25,
        steps: int = 5, warmup: int = 2):
    import torch.distributed as dist

    from .retrieval import ShardedIndexFlatL2, shard_bounds
    from .synthetic import planted_retrieval_set

    n_db = int(os.environ.get("NVS_RETR_NDB", 1_000_000)) if n_db is None else n_db
    n_q = int(os.environ.get("NVS_RETR_NQ", 10_000)) if n_q is None else n_q
    lo, hi = shard_bounds(n_db, world, rank)
    # every rank generates the identical global set chunk by chunk and keeps its own rows only
    db, q, planted = planted_retrieval_set(n_db, n_q, dim, k, seed=0, device=dev)
    shard = db if world == 1 else db[lo:hi].clone()
    del db
    torch.cuda.empty_cache()
    index = ShardedIndexFlatL2(dim, n_db, device=dev)
    index.add_local(shard)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        D, I = index.search(q, k)
    barrier()
    exact = bool(torch.equal(I, planted))
    gemm_ms, evs = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if world > 1:
            ge = []  # one event pair per pipelined query chunk
            index.search(q, k, gemm_events=ge)
            evs.append(ge)
        else:
            ge = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            index.search(q, k, gemm_events=ge)
            evs.append([ge])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    gemm = statistics.mean(sum(a.elapsed_time(b) for a, b in pairs) for pairs in evs)
    if world > 1:
        t = torch.tensor([ms, gemm], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, gemm = float(t[0]), float(t[1])
    flops_rank = 2.0 * n_q * (hi - lo) * dim
    return {
        "metric": "NetVLAD VPR queries/s", "value": n_q / (ms / 1e3), "unit": "queries/s",
        "config": {"workload": f"{n_q} queries x {n_db} db rows x {dim}-d, top-{k}, {world} shard(s) of {hi - lo} rows, "
                               "fp16 tcgen05 GEMM (exact screen) + fused lists + fp32 re-rank" +
                               (", global k-th bound exchange + NCCL all_gather merge, pipelined over query chunks" if world > 1 else "")},
        "ms_per_search": ms, "gemm_kernel_ms": gemm, "topk_bit_exact_vs_planted": exact, "scaling": "strong",
        "roofline": {"bound": "tensor", "achieved": flops_rank / (gemm / 1e3) / 1e12, "unit": "TFLOP/s",
                     "kind": "fp16 tcgen05.mma cta_group::2 256x256x16, clusters of 8 CTAs"},
    }


if __name__ == "__main__":
    # python -m nano_vs_slam_b200.retrieval_bench [n_db] [n_q]; under torchrun: one rank per GPU, sharded database
    import json
    import sys

    n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    r = run(dev, world, rank, n_db=n_db, n_q=n_q, steps=5 if world > 1 else 2, warmup=2 if world > 1 else 1)
    if rank == 0:
        print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()
