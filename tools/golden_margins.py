"""Print the parity margins (rel err vs the reference golden vectors) of every golden config, several runs."""
import contextlib, io, os, sys, glob
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import build_model, golden_cases, load_golden, rel_err
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
for path in golden_cases():
    c = load_golden(path)
    m = build_model(c["letter"], c["n_classes"], c["v3"], c["depth"], c["to_mcu"])
    m.load_state_dict(spread_init(m.state_dict(), c["wseed"])); m.eval(); m.training = False; m = m.cuda()
    x = synthetic_frames(c["B"], c["H"], c["W"], c["xseed"]).cuda()
    worst = {}
    for rep in range(int(os.environ.get("NVS_MARGIN_REPS", "20"))):
        out = m(x)
        for k in ("score", "coord", "feat", "vlad", "seg"):
            worst[k] = max(worst.get(k, 0.0), rel_err(out[k], c["fwd"][k]))
    print(os.path.basename(path), m.conv_backend, {k: f"{v:.1e}" for k, v in worst.items()})

# KITTI-size frame against the oracle (the case with the thinnest margin: sampled unit descriptors)
from oracle import kp2dtiny_ref as R
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory("S", 19, v3=False)
sd = spread_init(m.state_dict(), 1234)
m.load_state_dict(sd); m.eval(); m.training = False; m = m.cuda()
x = synthetic_frames(1, 376, 1241, 3)
a = R.arch_for("S", False, 19)
ref = R.forward(x, sd, a)
rpost = R.post_processing(dict(ref), 376, 1241, a)
worst = {}
for rep in range(5):
    out = m(x.cuda())
    post = m.post_processing(dict(out), 376, 1241)
    for k in ("score", "coord", "feat", "vlad", "seg"):
        worst[k] = max(worst.get(k, 0.0), rel_err(out[k], ref[k]))
    worst["post_feat"] = max(worst.get("post_feat", 0.0), rel_err(post["feat"], rpost["feat"]))
print("KITTI 376x1241 V2-S vs oracle", m.conv_backend, {k: f"{v:.2e}" for k, v in worst.items()})
