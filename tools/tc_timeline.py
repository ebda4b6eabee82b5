"""Per-step timeline of CTA 0 of conv_tc_kernel (debug build, -DNVS_TC_DEBUG): MMA issuers + one converter warp."""
import ctypes as C, os, sys, torch, numpy as np
NIS = int(os.environ.get("NVS_TC_ISSUERS", "2"))
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import _cabi
_cabi.LIB_PATH = "/root/repo/tools/libnanovs_dbg.so"
from nano_vs_slam_b200 import ops
lib = _cabi.lib()
lib.nvs_conv_tc_set_debug.argtypes = [C.c_void_p]
dbg = torch.zeros(4096 + 64 * 160, dtype=torch.int64, device="cuda")
lib.nvs_conv_tc_set_debug(dbg.data_ptr())
lib.nvs_conv_tc_set_knock.argtypes = [C.c_int]
lib.nvs_conv_tc_set_knock(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
cin, cout, H, W, B = 64, 64, 60, 80, 64
x = torch.randn(B, H, W, cin, device="cuda")
w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
out = torch.zeros(B, H, W, cout, device="cuda")
op = ops.TcConv(x, ops.pack_conv_tc(w, bias=b, math="tf32"), cout, act=1, dst=out)
for _ in range(3):
    dbg.zero_(); op.run()
torch.cuda.synchronize()
d = dbg.cpu().numpy()
m = d[:2048].reshape(512, 4); c = d[2048:4096].reshape(512, 4)
print("issuers (even steps: issuer 0, odd: issuer 1): step, wait sfull, issue 8 MMAs, commit, gap to own next step")
for i in range(36, 76):
    print(f"{i:4d} i{i%NIS}  wait {m[i,1]-m[i,0]:6d}  issue {m[i,2]-m[i,1]:6d}  commit {m[i,3]-m[i,2]:5d}   start-to-start(global) {m[i,0]-m[i-1,0]:6d}  own period {m[i,0]-m[i-NIS,0]:6d}")
per = np.diff(np.sort(m[18:324, 0])); print("mean global step period", per.mean(), "median", np.median(per), "total", m[323, 3] - m[0, 0])
print("converter warps 4 / 8 (groups 0 / 1): step, wait hfull, wait sempty, work")
idx = [i for i in range(512) if c[i, 0] != 0][18:40]
for i in idx:
    print(f"{i:4d}  hfull {c[i,1]-c[i,0]:6d}  sempty {c[i,2]-c[i,1]:6d}  work {c[i,3]-c[i,2]:6d}   sfull-arrive -> issuer sees it {m[i,1]-c[i,3]:6d}   issuer commit issued -> converter of step+4 passes sempty {c[i+4,2]-m[i,2] if i + 4 < 480 else 0:6d}")
