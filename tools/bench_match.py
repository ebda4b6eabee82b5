"""Matcher timing at VO size (4000 x 4000 x 32-d, ratio 0.7 one-to-one) and mutual-NN (HPatches, 1000 x 1000)."""
import sys, torch, torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import ops
g = torch.Generator().manual_seed(0)
for n, mode in ((4000, 0), (4000, 1), (1000, 1)):
    b = F.normalize(torch.randn(n, 32, generator=g), dim=1).cuda()
    a = F.normalize(b + 0.2 * torch.randn(n, 32, generator=g).cuda(), dim=1)
    for _ in range(3):
        ops.match(a, b, ratio=0.7, mode=mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        r = ops.match(a, b, ratio=0.7, mode=mode)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} mode={mode}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per pair, matches {int(r[3])}")
