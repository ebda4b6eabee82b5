"""Phase timing of the sharded retrieval search (torchrun, one rank per GPU): python -m torch.distributed.run ... tools/retr_tail.py
Every phase is bracketed by CUDA events on the search stream; the table is rank 0's mean over the timed searches."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200.retrieval import ShardedIndexFlatL2, merge_bounds, shard_bounds
from nano_vs_slam_b200.synthetic import planted_retrieval_set

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n_db, n_q, dim, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 10000, 4096, 25
lo, hi = shard_bounds(n_db, world, rank)
db, q, planted = planted_retrieval_set(n_db, n_q, dim, k, seed=0, device=dev)
shard = db[lo:hi].clone()
del db
torch.cuda.empty_cache()
idx = ShardedIndexFlatL2(dim, n_db, device=dev)
idx.add_local(shard)
ix = idx._index
names = ["begin (convert, GEMM, k-th values)", "all_gather bounds", "merge_bounds", "end (select, re-rank, scan)",
         "all_gather parts", "merge"]
acc = [0.0] * len(names)
gemm = 0.0
iters = 5
for it in range(iters + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ge = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    mine_b = ix.search_begin(q, k, gemm_events=ge)
    ev[1].record()
    all_b = torch.empty(world, n_q, k, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_b, mine_b)
    ev[2].record()
    gb = merge_bounds(all_b)
    ev[3].record()
    words = n_q * k + (n_q * k + 1) // 2
    mine = torch.empty(1, words, dtype=torch.int64, device=dev)
    Dm, Im = idx._packed_views(mine, 1, n_q, k)
    ix.search_end(q, k, gb, id_offset=lo, out=(Dm[0], Im[0]))
    ev[4].record()
    gathered = torch.empty(world, words, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, mine)
    ev[5].record()
    Dg, Ig = idx._packed_views(gathered, world, n_q, k)
    D, I = idx._merge(Dg, Ig)
    ev[6].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i in range(len(names)):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) / iters
        gemm += ge[0].elapsed_time(ge[1]) / iters
if rank == 0:
    print(f"N={world}: {n_q} queries x {n_db} rows ({hi - lo} per shard); exact {bool(torch.equal(I, planted))}; total {sum(acc):.3f} ms, GEMM kernel {gemm:.3f} ms")
    for n, a in zip(names, acc):
        print(f"  {n:36s} {a:8.3f} ms")
dist.destroy_process_group()
