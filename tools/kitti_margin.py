"""Worst relative error of every forward output at 376x1241 (V2-S, 19 classes) vs the oracle, both conv backends, with
a diagnosis of the worst sampled-descriptor elements (post_processing 'feat'): which cell, how small the descriptor
norm before normalisation is there (the division amplifies the dense map's error by 1 / norm), and how far the
sampling coordinate is off."""
import contextlib, io, os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import post_feat_errors, rel_err
from nano_vs_slam_b200 import tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
from oracle import kp2dtiny_ref as R
H, W = 376, 1241
a = R.arch_for("S", False, 19)
for backend in ("tc", "ffma"):
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory("S", 19, v3=False)
    m.conv_backend = backend
    sd = spread_init(m.state_dict(), 1234)
    m.load_state_dict(sd); m.eval(); m.training = False; m = m.cuda()
    for seed in (3, 4):
        x = synthetic_frames(1, H, W, seed)
        ref = R.forward(x, sd, a)
        rpost = R.post_processing(dict(ref), H, W, a)
        out = m(x.cuda())
        post = m.post_processing(dict(out), H, W)
        e = {k: rel_err(out[k], ref[k]) for k in ("score", "coord", "feat", "vlad", "seg")}
        e["post_feat"] = rel_err(post["feat"], rpost["feat"])
        print(backend, seed, {k: f"{v:.2e}" for k, v in e.items()},
              "| cells with identical coord: err %.2e, excess over the displacement bound elsewhere %.2e, identical %.3f"
              % post_feat_errors(post, rpost))
        # un-normalised sampled descriptors of the reference: grid_sample at the reference coordinates
        cn = rpost["coord"].clone()
        cn[:, 0] = cn[:, 0] / ((W - 1) / 2.0) - 1.0
        cn[:, 1] = cn[:, 1] / ((H - 1) / 2.0) - 1.0
        raw = F.grid_sample(ref["feat"], cn.permute(0, 2, 3, 1), align_corners=True)
        nrm = raw.norm(dim=1).flatten()
        d = (post["feat"].cpu() - rpost["feat"]).abs().amax(dim=1).flatten()
        top = d.topk(5)
        dc = (post["coord"].cpu() - rpost["coord"]).abs().amax(dim=1).flatten()
        print("   worst cells:", [(int(i), f"err {float(v):.2e}", f"norm {float(nrm[i]):.3f}", f"dcoord {float(dc[i]):.1e}")
                                   for v, i in zip(top.values, top.indices)],
              "| median norm %.3f, min norm %.3f" % (float(nrm.median()), float(nrm.min())))
        # the same error with the amplification taken out: |a - b| * norm / max|raw|
        print("   error x norm / max|raw|: %.2e" % float((d * nrm).max() / raw.abs().max()))
