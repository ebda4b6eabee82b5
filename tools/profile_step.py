"""Fixed launch sequence for ncu: W warm-up steps + 1 step of the bench workload (V2-S 240x320).
    ncu ... -k regex:conv_tc_kernel -s $((3*26+12)) -c 1 python tools/profile_step.py --batch 64
(26 conv_tc launches per step; #12 is the heaviest layer, desc_head.confAa 96->64 @120x160)."""
import argparse, contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import ops, tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--retrieval", action="store_true")
a = ap.parse_args()
if a.retrieval:
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    db, q, planted = planted_retrieval_set(262144, 4096, 4096, 25, seed=0, device="cuda")
    idx = IndexFlatL2(4096); idx.add(db)
    for _ in range(2):
        D, I = idx.search(q, 25)
    torch.cuda.synchronize()
    print("retrieval ok", bool(torch.equal(I, planted)))
    sys.exit(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory("S", 28, v3=False)
m.load_state_dict(spread_init(m.state_dict(), 1234)); m.eval(); m.training = False; m = m.cuda()
x = synthetic_frames(a.batch, 240, 320, 0).cuda()
for _ in range(a.warmup + 1):
    out = m(x); post = m.post_processing(out, 240, 320)
    sel = ops.select_keypoints(post["score"], post["coord"], post["feat"], 0.7, 1000)
torch.cuda.synchronize()
print("step ok", ops.LAUNCHES[0])
