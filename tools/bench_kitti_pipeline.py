"""Config 3: V2-S, 19 classes, 376x1241 frames, top-4000 keypoints (threshold 0: under spread-init weights no V2-S score passes 0.7, SURVEY allows top-k only), ratio-0.7 one-to-one
matching of consecutive frames.  Frames of a batch are consecutive (frame t+1 = frame t shifted + noise)."""
import contextlib, io, sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import ops, tiny_factory
from nano_vs_slam_b200.frontend import KP2DtinyFrontend
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
B, H, W = 32, 376, 1241
with contextlib.redirect_stdout(io.StringIO()):
    sd = spread_init(tiny_factory("S", 19).state_dict(), 1234)
    fe = KP2DtinyFrontend(config="S", nClasses=19, nn_thresh=0.0, top_k=4000, device="cuda", state_dict=sd)
x0 = synthetic_frames(1, H, W + 2 * B, 0)
g = torch.Generator().manual_seed(1)
x = torch.cat([x0[:, :, :, 2 * i:2 * i + W] + 0.01 * torch.randn(1, 3, H, W, generator=g) for i in range(B)]).cuda()

from nano_vs_slam_b200.matcher import match_consecutive

def step(match=True):
    sel, post = fe.run_batch(x, normalized=True)
    n_m = 0
    if match == "batched":  # one batched launch sequence, keypoint counts never leave the device
        i1, i2, dd, cnt = match_consecutive(sel)
        n_m = B - 1
    elif match:
        cnt = sel["count"].tolist()  # one small D2H per batch (the per-frame keypoint counts)
        for i in range(B - 1):
            a, b = sel["desc"][i, :cnt[i]], sel["desc"][i + 1, :cnt[i + 1]]
            if cnt[i] >= 1 and cnt[i + 1] >= 2:
                r = ops.match(a.contiguous(), b.contiguous(), ratio=0.7, mode=0)
                n_m += 1
    return sel, n_m

for m in (False, True, "batched"):
    for _ in range(3):
        step(m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        sel, n_m = step(m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"match={m}: {ms:.2f} ms per batch of {B} frames = {B / ms * 1e3:.0f} frames/s; keypoints/frame "
          f"{float(sel['count'].float().mean()):.0f}; pairs matched {n_m}")

sel, _ = step(False)
for _ in range(3):
    match_consecutive(sel)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    i1, i2, dd, cnt = match_consecutive(sel)
e1.record(); torch.cuda.synchronize()
print(f"match_consecutive alone: {e0.elapsed_time(e1) / 10:.3f} ms for {B - 1} pairs of 4000 keypoints; "
      f"matches/pair {float(cnt.float().mean()):.0f}")

# whole tail of the VO front-end: select -> batched matching -> batched relative pose, no host round trip
from nano_vs_slam_b200.matcher import pose_consecutive

K = (718.856, 718.856, 607.19, 185.22)  # KITTI sequence 00 pinhole intrinsics
for refine in (0, 10):
    for _ in range(3):
        pose_consecutive(sel, K, refine=refine)
    e0.record()
    for _ in range(10):
        (_, _, _, cnt), pose = pose_consecutive(sel, K, refine=refine)
    e1.record(); torch.cuda.synchronize()
    print(f"match + pose (refine {refine}): {e0.elapsed_time(e1) / 10:.3f} ms for {B - 1} pairs; "
          f"inliers/pair {float(pose['inliers'].float().mean()):.0f} of {float(cnt.float().mean()):.0f} matches")
