"""Generate tests/golden/*.npz by running the REAL reference package (build container only).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Imports ``src.kp2dtiny`` from /root/reference (read-only, never copied), loads name-hashed
spread-init weights (nano_vs_slam_b200.synthetic.spread_init) into the reference modules, runs
``model(x)`` + ``model.post_processing`` exactly as the reference callers do
(eval_multitask.py:195-198, frontend.py:56-60,84-85) and stores inputs' seeds and outputs.
Also stores cv2.BFMatcher 2-NN results for the matcher oracle.
The fixtures travel to the GPU box; /root/reference does not.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NVS_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)

from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames  # noqa: E402

# (name, letter, v3, n_classes, B, H, W, weight seed, input seed)
CASES = [
    ("v2_S", "S", False, 28, 2, 40, 56, 1234, 0),
    ("v2_S_A", "S_A", False, 19, 1, 40, 56, 1235, 1),
    ("v2_N", "N", False, 28, 1, 40, 56, 1236, 2),
    ("v2_N_A", "N_A", False, 19, 1, 32, 48, 1237, 3),
    ("v3_S", "S", True, 19, 1, 40, 56, 1238, 4),
    ("v3_S_A", "S_A", True, 28, 1, 32, 48, 1239, 5),
    ("v3_N", "N", True, 28, 2, 40, 56, 1240, 6),
    ("v3_N_A", "N_A", True, 19, 1, 40, 56, 1241, 7),
    # alternative aggregators (vpr.py:70-76); GeM's PixelUnshuffle(4) needs H/4 and W/4 to be multiples of 4
    ("v2_GEM_N", "GEM_N", False, 28, 2, 64, 80, 1242, 8),
    ("v2_GEM_S_A", "GEM_S_A", False, 19, 1, 48, 64, 1243, 9),
    ("v2_CONVAP_S_A", "CONVAP_S_A", False, 19, 1, 40, 56, 1244, 10),
    ("v3_CONVAP_S_A", "CONVAP_S_A", True, 28, 2, 40, 56, 1245, 11),
    # the large letter without attention (LARGE_D_V3: 64..512 channels, 128-d descriptors, ConvAP)
    ("v3_D", "D", True, 19, 1, 40, 56, 1246, 12),
    # the large attention letters (head_dim 64): V2 "D" and V3 "D_A"
    ("v2_D", "D", False, 19, 1, 32, 48, 1254, 20),
    ("v3_D_A", "D_A", True, 19, 1, 32, 48, 1255, 21),
    # cell 8 (TINY_F: downsample = 3, 16..256 channels, 64-d descriptors): skip-level map = H/4 must be a multiple of 4
    ("v2_F", "F", False, 19, 1, 64, 96, 1247, 13),
    # depth=True constructor kwarg: V2 second segmentation head (kp2dtiny.py:402-437), V3 middle slice + featD
    ("v2_S_depth", "S", False, 19, 1, 40, 56, 1248, 14),
    ("v3_N_depth", "N", True, 28, 2, 40, 56, 1249, 15),
    ("v3_S_A_depth", "S_A", True, 19, 1, 32, 48, 1250, 16),
    # to_mcu=True: ConvTranspose upsampling + ReLU (kp2dtiny.py:271-274).  KEEP THESE LAST: the reference's
    # get_config mutates its shared config dicts, every later model of the same letter would be an MCU model too.
    ("v2_S_mcu", "S", False, 19, 1, 40, 56, 1251, 17),
    ("v3_N_mcu", "N", True, 28, 1, 40, 56, 1252, 18),
    ("v2_S_A_mcu", "S_A", False, 19, 1, 32, 48, 1253, 19),
]


def load_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    sys.dont_write_bytecode = True
    from src.kp2dtiny.models.kp2dtiny import KP2DTinyV2, KP2DTinyV3, get_config, tiny_factory  # type: ignore

    def factory(letter, n_classes, v3=False, depth=False, to_mcu=False):
        if to_mcu:
            return tiny_factory(letter, n_classes, to_mcu=True, v3=v3)
        if not depth:
            return tiny_factory(letter, n_classes, v3=v3)
        cls = KP2DTinyV3 if v3 else KP2DTinyV2  # the callers' form: Cls(**get_config(...), nClasses=n, depth=True)
        return cls(**dict(get_config(letter, v3=v3)), nClasses=n_classes, depth=True)

    return factory


def run_reference(tiny_factory, letter, v3, n_classes, B, H, W, wseed, xseed, depth=False, to_mcu=False):
    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(letter, n_classes, v3=v3, depth=depth, to_mcu=to_mcu)
    sd = spread_init(m.state_dict(), wseed)
    m.load_state_dict(sd, strict=True)
    m.eval()
    m.training = False
    x = synthetic_frames(B, H, W, xseed)
    with torch.no_grad():
        out = m(x)
        fwd = {k: v.clone() for k, v in out.items()}
        post = m.post_processing(dict(out), H, W)
    return fwd, post


def main():
    tiny_factory = load_reference()
    gold = os.path.join(REPO, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    for name, letter, v3, ncls, B, H, W, wseed, xseed in CASES:
        depth, mcu = name.endswith("_depth"), name.endswith("_mcu")
        fwd, post = run_reference(tiny_factory, letter, v3, ncls, B, H, W, wseed, xseed, depth=depth, to_mcu=mcu)
        arrs = {"meta": np.array([int(v3), ncls, B, H, W, wseed, xseed], dtype=np.int64)}
        if depth:
            arrs["depth"] = np.array(1)
        if mcu:
            arrs["to_mcu"] = np.array(1)
        arrs["letter"] = np.array(letter)
        for k, v in fwd.items():
            arrs["fwd_" + k] = v.numpy()
        for k, v in post.items():
            arrs["post_" + k] = v.numpy()
        np.savez_compressed(os.path.join(gold, f"model_{name}.npz"), **arrs)
        print(name, {k: tuple(v.shape) for k, v in fwd.items()},
              "kp>0.7:", int((post["score"] > 0.7).sum()))

    # matcher: cv2.BFMatcher(NORM_L2).knnMatch(k=2) on unit descriptors with planted matches
    import cv2

    rng = np.random.default_rng(7)
    n1, n2, d = 300, 280, 32
    des2 = rng.standard_normal((n2, d)).astype(np.float32)
    des2 /= np.linalg.norm(des2, axis=1, keepdims=True)
    src = rng.integers(0, n2, size=n1)
    des1 = des2[src] + 0.25 * rng.standard_normal((n1, d)).astype(np.float32) * rng.random((n1, 1)).astype(np.float32)
    des1 /= np.linalg.norm(des1, axis=1, keepdims=True)
    des1 = des1.astype(np.float32)
    knn = cv2.BFMatcher(cv2.NORM_L2).knnMatch(des1, des2, k=2)
    idx = np.array([[m.trainIdx, n.trainIdx] for m, n in knn], dtype=np.int64)
    dist = np.array([[m.distance, n.distance] for m, n in knn], dtype=np.float32)
    cc = cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match(des1, des2)
    cross = np.array([[m.queryIdx, m.trainIdx] for m in cc], dtype=np.int64)
    np.savez_compressed(os.path.join(gold, "matcher_cv2.npz"), des1=des1, des2=des2, idx=idx, dist=dist,
                        cross=cross, cv2_version=np.array(cv2.__version__))
    print("matcher", idx.shape, cross.shape)


if __name__ == "__main__":
    main()
