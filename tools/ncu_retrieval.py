"""One retrieval search for `ncu --set full -k regex:flat_l2_topk` (10k queries x 262144 rows x 4096, top-25)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200.retrieval import IndexFlatL2
from nano_vs_slam_b200.synthetic import planted_retrieval_set
n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
db, q, planted = planted_retrieval_set(n_db, n_q, 4096, 25, seed=0, device="cuda")
ix = IndexFlatL2(4096)
ix.add(db)
D, I = ix.search(q, 25)
torch.cuda.synchronize()
print("exact", bool(torch.equal(I, planted)))
