"""Bottleneck experiments on conv_tc_kernel (debug build): kernel time with pipeline stages knocked out.
bits: 1 no MMAs, 2 converters only pass barriers, 4 no epilogue, 8 no tcgen05.st, 16 no halo LDS, 32 no weight TMA,
64 no halo TMA."""
import ctypes as C, sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import _cabi
_cabi.LIB_PATH = "/root/repo/tools/libnanovs_dbg.so"
from nano_vs_slam_b200 import ops
lib = _cabi.lib()
lib.nvs_conv_tc_set_debug.argtypes = [C.c_void_p]
lib.nvs_conv_tc_set_knock.argtypes = [C.c_int]
dbg = torch.zeros(4096 + 64 * 160, dtype=torch.int64).pin_memory()
lib.nvs_conv_tc_set_debug(dbg.data_ptr())
lib.nvs_conv_tc_set_timeout_log.argtypes = [C.c_void_p]
tlog = torch.zeros(1 + 4 * 200, dtype=torch.int64).pin_memory()  # host-mapped: readable after a device fault
lib.nvs_conv_tc_set_timeout_log(tlog.data_ptr())
NH, NS = 3, 4
BAR_NAMES = [f"hfull{i}" for i in range(NH)] + [f"hempty{i}" for i in range(NH)] + [f"sfull{i}" for i in range(NS)] + \
    [f"sempty{i}" for i in range(NS)] + ["afull0", "afull1", "aempty0", "aempty1", "astart0", "astart1"]
def dump_timeouts(fault=False):
    n = int(tlog[0])
    if not n:
        return
    r = tlog[1:1 + 4 * min(n, 200)].numpy().copy().reshape(-1, 4)
    r = r[r[:, 0].argsort()]
    print(f"  {n} mbarrier timeouts; first per warp (bar index relative to the lowest address seen):")
    base = min(int(v) >> 8 for v in r[:, 2])
    seen = set()
    for t0, bt, bp, line in r:
        idx = ((int(bp) >> 8) - base) // 8
        key = (int(bt) >> 32, (int(bt) & 0xffffffff) // 32)
        if key in seen:
            continue
        seen.add(key)
        print(f"    first timeout: t0 {t0 - r[0, 0]:10d}  block {key[0]:4d} warp {key[1]:2d}  bar +{idx:3d}  parity {int(bp) & 1}  line {line}")
    tlog.zero_()
cases = [(64, 64, 60, 80, 256), (64, 128, 60, 80, 256), (16, 32, 240, 320, 64)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in sys.argv[1].split(","))]
for cin, cout, H, W, B in cases:
    x = torch.randn(B, H, W, cin, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    out = torch.zeros(B, H, W, cout, device="cuda")
    op = ops.TcConv(x, ops.pack_conv_tc(w, bias=b, math="tf32"), cout, act=1, dst=out)
    steps = 9 * ((cin + 31) // 32) if cin >= 32 else 9
    tiles = B * ((H + 7) // 8) * ((W + 15) // 16)
    print(f"== {cin}->{cout} @{H}x{W} B={B}: {tiles} tiles, {steps} steps/tile")
    variants = [(int(sys.argv[2]), "knock " + sys.argv[2])] if len(sys.argv) > 2 else None
    for bits, name in variants or [(0, "full"), (1, "no MMA"), (2, "no converter work"), (4, "no epilogue"), (8, "no tcgen05.st"),
                       (16, "no halo LDS"), (3, "no MMA, no converter"), (7, "barriers only"), (5, "converter only"), (6, "MMA only"),
                       (32, "no W TMA"), (64, "no halo TMA"), (96, "no TMA"), (7 + 96, "barriers only, no TMA"), (6 + 96, "MMA only, no TMA"),
                       (5 + 96, "converter only, no TMA"), (6 + 32, "MMA only, no W TMA"), (4 + 96, "no epilogue, no TMA"), (16 + 96, "no LDS, no TMA")]:
        lib.nvs_conv_tc_set_knock(bits)
        op.run()
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("  device fault:", str(e).splitlines()[0]); dump_timeouts(True); sys.exit(1)
        dump_timeouts()
        for _ in range(2):
            op.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            op.run()
        e1.record()
        try:
            torch.cuda.synchronize()
        except Exception as e:
            print("  device fault:", str(e).splitlines()[0]); dump_timeouts(True); sys.exit(1)
        dump_timeouts()
        ms = e0.elapsed_time(e1) / 10
        per_step = ms * 1e-3 * 1.9e9 / (tiles / 148) / steps
        print(f"  {name:24s} {ms*1e3:8.1f} us   ~{per_step:6.0f} cycles/step @1.9GHz")
    lib.nvs_conv_tc_set_knock(0)
