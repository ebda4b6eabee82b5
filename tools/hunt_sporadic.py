"""Repeat one forward many times and report every run whose outputs deviate from the first run by more than the
multi-issuer rounding noise (a protocol race would show up as rare outliers)."""
import contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
letter, v3, B, H, W, n = sys.argv[1], sys.argv[2] == "1", int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory(letter, 19, v3=v3)
m.load_state_dict(spread_init(m.state_dict(), 1234)); m.eval(); m.training = False; m = m.cuda()
m.cuda_graph_max_batch = int(os.environ.get("GRAPH", "16"))
x = synthetic_frames(B, H, W, 3).cuda()
ref = {k: v.clone() for k, v in m(x).items()}
worst = {k: 0.0 for k in ref}
bad = 0
for i in range(n):
    out = m(x)
    for k in ref:
        e = float((out[k] - ref[k]).abs().max() / ref[k].abs().max())
        worst[k] = max(worst[k], e)
        if e > 1e-5:
            bad += 1
            print("outlier run", i, k, f"{e:.2e}")
print(letter, B, H, W, "runs", n, "outliers", bad, {k: f"{v:.1e}" for k, v in worst.items()})
