"""Per-step timeline of the MMA-issuing thread of CTA 0 of conv_tc_kernel (debug build, -DNVS_TC_DEBUG)."""
import ctypes as C, sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import _cabi
_cabi.LIB_PATH = "/root/repo/tools/libnanovs_dbg.so"
from nano_vs_slam_b200 import ops
lib = _cabi.lib()
lib.nvs_conv_tc_set_debug.argtypes = [C.c_void_p]
dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.nvs_conv_tc_set_debug(dbg.data_ptr())
cin, cout, H, W, B = 64, 64, 60, 80, 64
x = torch.randn(B, H, W, cin, device="cuda")
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
out = torch.zeros(B, H, W, cout, device="cuda")
op = ops.TcConv(x, ops.pack_conv_tc(w, bias=b), cout, act=1, dst=out)
for _ in range(3):
    dbg.zero_(); op.run()
torch.cuda.synchronize()
m = dbg.cpu().numpy()[:2048].reshape(512, 4)
print("step: issue6(first MMAs) | poll next barriers | issue2+commit | period")
for i in range(36, 76):
    print(f"{i:4d}  first {m[i,1]-m[i,0]:6d}  poll {m[i,2]-m[i,1]:6d}  rest {m[i,3]-m[i,2]:5d}   period {m[i,0]-m[i-1,0]:6d}")
per = np.diff(m[18:324, 0])
print("mean step period", per.mean(), "median", np.median(per), "total", m[323, 3] - m[0, 0])
