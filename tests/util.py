"""Shared helpers for the parity tests (tolerances from BASELINE.json:north_star)."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-4  # fp32 features / descriptors / VLAD: 1e-4 relative (north_star)


def tol(key: str, v3: bool) -> float:
    """Tolerance per forward output: 1e-4 relative for every tensor on every backend (north_star).  V3 returns Softmax2d
    probabilities under 'seg' (kp2dtiny.py:942-943); they are held to the same 1e-4 (measured worst case of the default
    tensor-core backend over the golden configs: 5e-5, profiles/r2_parity_margins_slice128.txt)."""
    return REL_TOL


# Sampled + L2-normalised descriptors (post_processing 'feat') are a function of the decoded keypoint coordinate, and
# that coordinate is an fp32 number of up to ~1240 px: one ulp is 1.2e-4 px beyond x = 1024.  A shift difference of a few
# 1e-6 (well inside every tolerance) can round the coordinate to the neighbouring fp32 value, the descriptor is then
# sampled 6e-5 feature-map pixels away, and that alone moves its components by ~1e-4 -- on ANY backend (measured:
# exact-fp32 FFMA backend 1.15e-4 / 1.27e-4, tensor-core backend 1.13e-4 / 1.23e-4 at 376x1241, every worst cell with a
# coordinate difference of exactly one ulp; tools/kitti_margin.py).  So descriptors are compared where the coordinate
# is bit-identical (1e-4, the north_star bound), and cells whose coordinate differs (by <= 1e-3 px, the north_star
# coordinate tolerance) are bounded by what that displacement can do: 1e-4 + |d coord| in feature-map pixels x the largest
# neighbour-to-neighbour step of the reference descriptor map.
def post_feat_errors(post: dict, rpost: dict):
    """-> (relative error over cells with bit-identical coordinates, worst excess over the displacement bound elsewhere,
    fraction of cells with identical coordinates)."""
    f, rf = post["feat"].detach().float().cpu(), rpost["feat"].detach().float().cpu()
    c, rc = post["coord"].detach().float().cpu(), rpost["coord"].detach().float().cpu()
    dc = (c - rc).abs().amax(dim=1)                    # (B, Hc, Wc) pixels
    same = dc == 0
    scale = rf.abs().max().clamp_min(1e-30)
    err = (f - rf).abs().amax(dim=1) / scale           # (B, Hc, Wc)
    e_same = float(err[same].max()) if bool(same.any()) else 0.0
    # neighbour-to-neighbour variation of the reference's unit descriptors (cells are 4 px apart; the sampled map has
    # half the image resolution): a conservative per-feature-pixel slope
    gx = (rf[..., :, 1:] - rf[..., :, :-1]).abs().max() if rf.shape[-1] > 1 else rf.new_zeros(())
    gy = (rf[..., 1:, :] - rf[..., :-1, :]).abs().max() if rf.shape[-2] > 1 else rf.new_zeros(())
    slope = float(torch.maximum(gx, gy) / scale)
    bound = 1e-4 + 0.5 * dc * slope
    excess = float((err - bound)[~same].max()) if bool((~same).any()) else -1.0
    return e_same, excess, float(same.float().mean())


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_inf / ||b||_inf."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz")))


def load_golden(path):
    z = np.load(path, allow_pickle=False)
    v3, ncls, B, H, W, wseed, xseed = [int(v) for v in z["meta"]]
    letter = str(z["letter"])
    fwd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fwd_")}
    post = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("post_")}
    return dict(letter=letter, v3=bool(v3), n_classes=ncls, B=B, H=H, W=W, wseed=wseed, xseed=xseed,
                fwd=fwd, post=post, depth="depth" in z.files, to_mcu="to_mcu" in z.files)


def argmax_agreement(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.cpu() == b.cpu()).float().mean())


def build_model(letter, n_classes, v3=False, depth=False, to_mcu=False):
    """tiny_factory, or the callers' explicit form when depth=True (Cls(**get_config(...), nClasses=n, depth=True))."""
    import contextlib
    import io

    from nano_vs_slam_b200 import KP2DTinyV2, KP2DTinyV3, get_config, tiny_factory

    with contextlib.redirect_stdout(io.StringIO()):
        if to_mcu:
            return tiny_factory(letter, n_classes, to_mcu=True, v3=v3)
        if not depth:
            return tiny_factory(letter, n_classes, v3=v3)
        return (KP2DTinyV3 if v3 else KP2DTinyV2)(**get_config(letter, v3=v3), nClasses=n_classes, depth=True)
