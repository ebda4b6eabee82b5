import contextlib, io, sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
def mk(backend):
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory("S", 19, v3=False)
    m.conv_backend = backend
    m.cuda_graph_max_batch = 0
    m.load_state_dict(spread_init(m.state_dict(), 4321)); m.eval(); m.training = False
    return m.cuda()
H, W = 376, 1241
x = synthetic_frames(1, H, W, 17).cuda()
b = mk("ffma"); ob = b(x); pb = {k: v.clone() for k, v in next(iter(b._plans.values())).bufs.items()}
a = mk("tc")
names = ["t1a","p1","t2a","t2b","t3a","skip","p3","t4a","xb","sh","lh","da","dps","dA","s0","sp","s2","sp3","ps1","s5","ps2","s7","v1","v2","v3"]
for rep in range(30):
    oa = a(x)
    pa = next(iter(a._plans.values())).bufs
    bad = []
    for k in names:
        ta, tb = pa[k], pb[k]
        if ta.shape != tb.shape: ta = ta.permute(0, 3, 1, 2)
        err = float((ta - tb).abs().max() / tb.abs().max())
        if err > 1e-4:
            d = (ta - tb).abs()[0].amax(0)
            ys, xs = torch.nonzero(d > 1e-4 * tb.abs().max(), as_tuple=True)
            bad.append((k, f"{err:.2e}", int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max()), len(ys)))
    outs = {k: float((oa[k]-ob[k]).abs().max()/ob[k].abs().max()) for k in ("score","coord","feat","seg","vlad")}
    print(rep, "BAD" if bad else "ok", bad[:3], {k: f"{v:.1e}" for k, v in outs.items()} if bad else "")
