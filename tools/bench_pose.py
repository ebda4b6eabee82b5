"""Timing of nvs_pose_batch at the KITTI batch shape (31 consecutive pairs, up to 4000 matches each, 512 samples)."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import ops  # noqa: E402

P, kmax, iters = 31, 4000, 512
rng = np.random.default_rng(0)
X = np.stack([rng.uniform(-4, 4, (P, kmax)), rng.uniform(-2, 2, (P, kmax)), rng.uniform(4, 30, (P, kmax))], -1)
cur = X[..., :2] / X[..., 2:]
X2 = X + np.array([0.05, -0.02, 1.0])
ref = X2[..., :2] / X2[..., 2:] + rng.normal(0, 2e-4, cur.shape)
out = rng.random((P, kmax)) < 0.25
ref[out] = rng.uniform(-0.5, 0.5, (int(out.sum()), 2))
pts = torch.from_numpy(np.concatenate([cur, ref]).astype(np.float32)).cuda()
a = torch.arange(P, dtype=torch.int32, device="cuda")
cnt = torch.full((P,), kmax, dtype=torch.int32, device="cuda")
ws = torch.empty(int(ops.lib().nvs_pose_workspace_bytes(P, kmax, iters)), dtype=torch.uint8, device="cuda")
for refine in (0, 10):
    for _ in range(3):
        o = ops.pose_batch(pts, a, a + P, cnt, iters=iters, refine=refine, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        o = ops.pose_batch(pts, a, a + P, cnt, iters=iters, refine=refine, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    print("pose_batch %d pairs x %d matches x %d samples, refine %d: %.3f ms per call, inliers %s" %
          (P, kmax, iters, refine, e0.elapsed_time(e1) / 10, o["inliers"][:4].tolist()))
