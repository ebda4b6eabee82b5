"""Worst relative error of every forward output at 376x1241 (V2-S, 19 classes) over several runs, vs the oracle."""
import contextlib, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import rel_err
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
from oracle import kp2dtiny_ref as R
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory("S", 19, v3=False)
sd = spread_init(m.state_dict(), 1234)
m.load_state_dict(sd); m.eval(); m.training = False; m = m.cuda()
a = R.arch_for("S", False, 19)
for seed in (3, 4):
    x = synthetic_frames(1, 376, 1241, seed)
    ref = R.forward(x, sd, a)
    rpost = R.post_processing(dict(ref), 376, 1241, a)
    for rep in range(4):
        out = m(x.cuda())
        post = m.post_processing(dict(out), 376, 1241)
        e = {k: rel_err(out[k], ref[k]) for k in ("score", "coord", "feat", "vlad", "seg")}
        e["post_feat"] = rel_err(post["feat"], rpost["feat"])
        d = (out["coord"].cpu() - ref["coord"]).abs()
        i = int(d.argmax())
        print(seed, rep, {k: f"{v:.2e}" for k, v in e.items()}, "coord worst at", i, float(ref["coord"].view(-1)[i]))
