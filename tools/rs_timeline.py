"""Per-role clock64 stamps of CTA 0 of one conv_rs launch (NVS_RS_KNOCK applies): python tools/rs_timeline.py cin cout H W B"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import ops
from nano_vs_slam_b200._cabi import lib
cin, cout, H, W, B = (int(v) for v in sys.argv[1:6]) if len(sys.argv) > 5 else (32, 32, 120, 160, 256)
x = torch.randn(B, H, W, cin, device="cuda")
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
out = torch.zeros(B, H, W, (cout + 31) // 32 * 32, device="cuda")
op = ops.tc_conv(x, ops.pack_conv_tc(w, bias=b, math="f16"), cout, act=1, dst=out)
for _ in range(3):
    op.run()
buf = torch.zeros(768, dtype=torch.int64, device="cuda")
lib().nvs_conv_rs_debug_buffer(buf.data_ptr())
op.run()
torch.cuda.synchronize()
lib().nvs_conv_rs_debug_buffer(None)
t = buf.cpu().numpy()
for name, off in (("epilogue/tile", 0), ("converter0/row", 256), ("mma/chunk", 512)):
    v = t[off:off + 256]
    v = v[v > 0]
    d = v[1:] - v[:-1]
    print(f"{name:16s} n={len(v)} median {int(sorted(d)[len(d)//2]) if len(d) else 0} mean {d.mean() if len(d) else 0:.0f}  first 24 deltas: {d[:24].tolist()}")
