"""One pass over the kernels that are NOT the two tcgen05 contractions, for `ncu --set full` (VERDICT r1 row g):

    python tools/ncu_workload.py small   # stem, netvlad, decode, seg_argmax, select at V2-S 240x320 batch 64;
                                         # knn2 / one_to_one at 4000 x 4000 x 32; the pose kernels at 31 x 4000 x 512
    python tools/ncu_workload.py att     # attention / LN / dw3x3 / 1x1 of V2-S_A at 512x1024 (BASELINE config 4), batch 2

Every kernel of interest launches exactly once (ncu replays each captured launch ~40 times)."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import ops, tiny_factory  # noqa: E402
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "small"


def model(letter, ncls):
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(letter, ncls, v3=False)
    m.load_state_dict(spread_init(m.state_dict(), 1234))
    m.eval()
    m.training = False
    m = m.cuda()
    m.cuda_graph_max_batch = 0
    return m


if what == "att":
    m = model("S_A", 19)
    x = synthetic_frames(2, 512, 1024, 0).cuda()
    out = m(x)
    torch.cuda.synchronize()
    print("att pass done", tuple(out["seg"].shape))
else:
    m = model("S", 28)
    x = synthetic_frames(64, 240, 320, 0).cuda()
    out = m(x)
    post = m.post_processing(out, 240, 320)
    sel = ops.select_keypoints(post["score"], post["coord"], post["feat"], 0.7, 1000)
    g = torch.Generator().manual_seed(0)
    b = F.normalize(torch.randn(4000, 32, generator=g), dim=1).cuda()
    a = F.normalize(b + 0.2 * torch.randn(4000, 32, generator=g).cuda(), dim=1)
    r = ops.match(a, b, ratio=0.7, mode=0)
    P, kmax, iters = 31, 4000, 512
    rng = np.random.default_rng(0)
    X = np.stack([rng.uniform(-4, 4, (P, kmax)), rng.uniform(-2, 2, (P, kmax)), rng.uniform(4, 30, (P, kmax))], -1)
    cur = X[..., :2] / X[..., 2:]
    X2 = X + np.array([0.05, -0.02, 1.0])
    ref = X2[..., :2] / X2[..., 2:] + rng.normal(0, 2e-4, cur.shape)
    bad = rng.random((P, kmax)) < 0.25
    ref[bad] = rng.uniform(-0.5, 0.5, (int(bad.sum()), 2))
    pts = torch.from_numpy(np.concatenate([cur, ref]).astype(np.float32)).cuda()
    ia = torch.arange(P, dtype=torch.int32, device="cuda")
    cnt = torch.full((P,), kmax, dtype=torch.int32, device="cuda")
    o = ops.pose_batch(pts, ia, ia + P, cnt, iters=iters, refine=10)
    torch.cuda.synchronize()
    print("small pass done", int(sel["count"].sum()), int(r[3]), o["inliers"][:3].tolist())
