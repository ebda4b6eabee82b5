"""Print the parity margins (rel err vs the reference golden vectors) of every golden config, several runs."""
import contextlib, io, os, sys, glob
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import golden_cases, load_golden, rel_err
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
for path in golden_cases():
    c = load_golden(path)
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(c["letter"], c["n_classes"], v3=c["v3"])
    m.load_state_dict(spread_init(m.state_dict(), c["wseed"])); m.eval(); m.training = False; m = m.cuda()
    x = synthetic_frames(c["B"], c["H"], c["W"], c["xseed"]).cuda()
    worst = {}
    for rep in range(20):
        out = m(x)
        for k in ("score", "coord", "feat", "vlad", "seg"):
            worst[k] = max(worst.get(k, 0.0), rel_err(out[k], c["fwd"][k]))
    print(os.path.basename(path), m.conv_backend, {k: f"{v:.1e}" for k, v in worst.items()})
