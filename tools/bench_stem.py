"""Stem layer (3->16 @240x320) timing: dedicated kernel (NHWC store) vs the generic tiled kernel (NCHW store)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import ops
B = 256
x = (torch.rand(B, 3, 240, 320, device="cuda") * 2 - 1)
w = torch.randn(16, 3, 3, 3, device="cuda") * 0.3
b = torch.randn(16, device="cuda") * 0.1
wp, bp = ops.pack_conv(w, bias=b)
for nhwc in (True, False):
    for _ in range(3):
        ops.conv(x, wp, bp, 16, act=1, dst_nhwc=nhwc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv(x, wp, bp, 16, act=1, dst_nhwc=nhwc)
    e1.record(); torch.cuda.synchronize()
    print("nhwc" if nhwc else "nchw", e0.elapsed_time(e1) / 10, "ms")
