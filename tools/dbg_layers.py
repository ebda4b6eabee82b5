import contextlib, io, sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
def mk(backend):
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory("S", 19, v3=False)
    m.conv_backend = backend
    m.load_state_dict(spread_init(m.state_dict(), 4321)); m.eval(); m.training = False
    return m.cuda()
for (H, W) in ((376, 1241), (72, 153), (240, 320)):
    x = synthetic_frames(1, H, W, 17).cuda()
    a, b = mk("tc"), mk("ffma")
    oa, ob = a(x), b(x)
    pa = next(iter(a._plans.values())).bufs; pb = next(iter(b._plans.values())).bufs
    print("shape", H, W)
    for k in ["t1a","p1","t2a","t2b","t3a","skip","p3","t4a","xb","sh","lh","da","dps","dA","s0","sp","s2","sp3","ps1","s5","ps2","s7","v1","v2","v3"]:
        if k in pa and k in pb:
            ta, tb = pa[k], pb[k]
            if ta.shape != tb.shape: ta = ta.permute(0, 3, 1, 2)
            err = float((ta - tb).abs().max() / tb.abs().max())
            flag = "  <-- BAD" if err > 1e-4 else ""
            print(f"  {k:5s} {tuple(tb.shape)} rel {err:.2e}{flag}")
            if err > 1e-4:
                d = (ta - tb).abs()[0].amax(0)
                ys, xs = torch.nonzero(d > 1e-4 * tb.abs().max(), as_tuple=True)
                print("     bad rows", ys.min().item(), ys.max().item(), "cols", xs.min().item(), xs.max().item(), "count", len(ys))
    for k in ("score","coord","feat","seg","vlad"):
        print("  out", k, float((oa[k]-ob[k]).abs().max()/ob[k].abs().max()))
