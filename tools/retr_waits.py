"""Wait profile of the retrieval GEMM (nvs_flat_debug_buffer): where each role of flat_l2_topk_kernel spends its cycles.
python tools/retr_waits.py [n_db] [n_q]   (NVS_RETR_CLUSTER / NVS_RETR_STAGES apply)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200._cabi import lib
from nano_vs_slam_b200.retrieval import IndexFlatL2
from nano_vs_slam_b200.synthetic import planted_retrieval_set
n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
db, q, _ = planted_retrieval_set(n_db, n_q, 4096, 25, seed=0, device="cuda")
ix = IndexFlatL2(4096)
ix.add(db)
for _ in range(2):
    ix.search(q, 25)
buf = torch.zeros(16 * 148, dtype=torch.int64, device="cuda")
lib().nvs_flat_debug_buffer(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ix.search(q, 25)
e1.record()
torch.cuda.synchronize()
lib().nvs_flat_debug_buffer(None)
t = buf.cpu().view(148, 16).double()
te = t[t[:, 5] > 0]
t = t[t[:, 0] > 0]
tot = t[:, 0]
ms = e0.elapsed_time(e1)
print(f"search {ms:.2f} ms; {len(t)} CTAs; MMA-role cycles mean {tot.mean():.0f} max {tot.max():.0f} -> {tot.max() / ms / 1e3:.0f} MHz")
print(f"cycles per k-block: {(tot / t[:, 6]).mean():.0f} (512 = tensor pipe saturated)")
for name, c in (("producer waits for a free stage", 1), ("MMA waits for operands", 2), ("MMA waits for an accumulator", 3),
                ("epilogue waits for MMA", 4)):
    print(f"  {name:32s} {100 * (t[:, c] / tot).mean():5.1f} % of the kernel (min {100 * (t[:, c] / tot).min():.1f}, max {100 * (t[:, c] / tot).max():.1f})")
et = te[:, 5]
for name, c in (("epilogue: tcgen05.ld + wait::ld", 7), ("epilogue: named barrier", 8)):
    print(f"  {name:32s} {100 * (te[:, c] / et).mean():5.1f} % of the kernel (min {100 * (te[:, c] / et).min():.1f}, max {100 * (te[:, c] / et).max():.1f})")
print(f"  32-column chunks with an append per CTA: one row {te[:, 9].mean():.0f}, the 32 rows of a warp {te[:, 10].mean():.0f}; compactions of one row {te[:, 11].mean():.0f}")
