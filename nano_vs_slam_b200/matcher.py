"""BfFeatureMatcher -- device version of src/visual_odometry/feature_matcher.py:234-248 (+ :89-98, :179-209).

``match(des1, des2)`` accepts numpy arrays (as the reference does) or CUDA tensors and returns the
reference's ``(idx1, idx2, score)`` python lists.  ``match_device`` returns device tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, torch_ops  # noqa: F401  (torch_ops registers torch.ops.nanovs.*)

kRatioTest = 0.7  # feature_matcher.py:26
NORM_L2 = 4       # cv2.NORM_L2


class BfFeatureMatcher(object):
    def __init__(self, norm_type=NORM_L2, cross_check=False, ratio_test=kRatioTest, type=None, device="cuda"):
        if norm_type != NORM_L2:
            raise NotImplementedError("only NORM_L2 float descriptors are on the hot path (feature_matcher.py:248)")
        self.norm_type = norm_type
        self.cross_check = cross_check
        self.ratio_test = ratio_test
        self.device = device
        self.matcher_name = "BfFeatureMatcher"
        self.matches = None

    def _dev(self, d):
        if isinstance(d, np.ndarray):
            d = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
        return d.to(self.device, dtype=torch.float32).contiguous()

    def match_device(self, des1, des2, ratio_test=None):
        ratio = self.ratio_test if ratio_test is None else ratio_test
        mode = 1 if self.cross_check else 0
        return torch.ops.nanovs.match(self._dev(des1), self._dev(des2), float(ratio), mode)

    def knn_match(self, des1, des2):
        """cv2.BFMatcher.knnMatch(des1, des2, k=2) as (idx (n1,2), dist (n1,2)) device tensors."""
        idx, _, dist, _ = torch.ops.nanovs.match(self._dev(des1), self._dev(des2), kRatioTest, 2)
        return idx, dist

    def match(self, des1, des2, ratio_test=None):
        i1, i2, dd, cnt = self.match_device(des1, des2, ratio_test)
        n = int(cnt)
        return i1[:n].cpu().tolist(), i2[:n].cpu().tolist(), dd[:n].cpu().tolist()


@torch.no_grad()
def match_selected(sel: dict, i: int, j: int, cross_check: bool = True, ratio_test: float = kRatioTest):
    """Match the keypoints of frames ``i`` and ``j`` of ONE ``ops.select_keypoints`` result without leaving the
    device, and bring the matched coordinates back in a single D2H copy.

    This is the data flow of ``compute_homography`` / ``compute_matching_score`` (evaluation/descriptor.py:221-229:
    select_k_best -> BFMatcher(crossCheck=True).match -> keypoints[matches]) and of the VO loop
    (visual_odometry.py:314-322: BfFeatureMatcher.match -> kps[idx]) with the per-pair host round trips removed.
    Returns numpy (m,2), (m,2), (m,): matched points of frame i, of frame j, descriptor distances."""
    ni, nj = int(sel["count"][i]), int(sel["count"][j])
    if ni == 0 or nj == 0:
        z = np.zeros((0, 2), np.float32)
        return z, z.copy(), np.zeros((0,), np.float32)
    i1, i2, dd, cnt = torch.ops.nanovs.match(sel["desc"][i, :ni].contiguous(), sel["desc"][j, :nj].contiguous(),
                                             float(ratio_test), 1 if cross_check else 0)
    # one packed tensor = one D2H: rows [x_i, y_i, x_j, y_j, dist], then the count in the last row
    m = i1.shape[0]
    packed = torch.empty(m + 1, 5, device=i1.device, dtype=torch.float32)
    valid = torch.arange(m, device=i1.device) < cnt.reshape(-1)[0]  # entries past the count are unspecified
    z = torch.zeros_like(i1)
    packed[:m, 0:2] = sel["pts"][i][torch.where(valid, i1, z).long()]
    packed[:m, 2:4] = sel["pts"][j][torch.where(valid, i2, z).long()]
    packed[:m, 4] = dd
    packed[m] = cnt.to(torch.float32)
    host = packed.cpu().numpy()
    n = int(host[m, 0])
    return host[:n, 0:2].copy(), host[:n, 2:4].copy(), host[:n, 4].copy()


@torch.no_grad()
def lightglue_inputs(sel: dict, i: int, j: int, image_size):
    """LightGlue hand-off (visual_odometry.py:207-227, 236-249): keypoints divided by (W, H), descriptors and the
    image size as batch-1 device tensors, built from a select_keypoints result without a host round trip.
    ``image_size`` = (H, W) as ``new_size`` in the reference."""
    H, W = image_size
    dev = sel["pts"].device
    scale = torch.tensor([W, H], dtype=torch.float32, device=dev).unsqueeze(0)
    size = torch.tensor([[W, H]], dtype=torch.float32, device=dev)
    out = {}
    for tag, f in (("0", i), ("1", j)):
        n = int(sel["count"][f])
        out["image" + tag] = {"keypoints": (sel["pts"][f, :n] / scale).unsqueeze(0),
                              "descriptors": sel["desc"][f, :n].unsqueeze(0), "image_size": size}
    return out


@torch.no_grad()
def match_consecutive(sel: dict, cross_check: bool = False, ratio_test: float = kRatioTest):
    """VO data flow for a batch of consecutive frames (visual_odometry.py:314-322 matches frame t against t-1):
    every frame i+1 (query) against frame i (train) in ONE batched launch sequence, counts read on the device.
    Returns device tensors (idx_cur (B-1,k), idx_prev (B-1,k), dist (B-1,k), count (B-1,))."""
    B = sel["desc"].shape[0]
    dev = sel["desc"].device
    a = torch.arange(1, B, device=dev, dtype=torch.int32)
    return torch.ops.nanovs.match_batch(sel["desc"], sel["count"], a, a - 1, float(ratio_test), 1 if cross_check else 0)


def pose_consecutive(sel: dict, intrinsics, cross_check: bool = False, ratio_test: float = kRatioTest,
                     threshold: float = 0.0003, iters: int = 512, seed: int = 0, refine: int = 0,
                     confidence: float = 0.0):
    """Matching AND relative pose for a batch of consecutive frames without a host round trip
    (visual_odometry.py:314-345: match frame t against t-1, then estimatePose :383-412).

    ``intrinsics`` = (fx, fy, cx, cy).  Returns (matches, pose): ``matches`` as match_consecutive, ``pose`` =
    dict(E, R, t, mask, inliers) per pair with x_prev ~ R x_cur + t (the reference's kpn_cur -> kpn_ref order)."""
    B = sel["desc"].shape[0]
    dev = sel["desc"].device
    a = torch.arange(1, B, device=dev, dtype=torch.int32)
    i1, i2, dd, cnt = torch.ops.nanovs.match_batch(sel["desc"], sel["count"], a, a - 1, float(ratio_test),
                                                   1 if cross_check else 0)
    pose = torch_ops.pose_batch(sel["pts"], a, a - 1, cnt, i1, i2, intrinsics=intrinsics, threshold=threshold,
                                iters=iters, seed=seed, refine=refine, confidence=confidence)
    return (i1, i2, dd, cnt), pose


class PoseEstimator(object):
    """Reference-shaped pose step: ``estimatePose(kps_ref, kps_cur) -> (R, t)`` with ``mask_match`` set, as
    VisualOdometry.estimatePose (visual_odometry.py:383-412).  ``cam`` is any object with fx, fy, cx, cy (the
    reference's PinholeCamera); lens distortion is not handled here (KITTI frames are rectified: D = 0)."""

    def __init__(self, cam, threshold: float = 0.0003, iters: int = 512, seed: int = 0, refine: int = 0,
                 device: str = "cuda", prob: float = 0.0):
        """``prob`` > 0 (the reference passes 0.999, visual_odometry.py:392): stop sampling once the RANSAC confidence
        bound for the best consensus so far is reached (at most ``iters`` samples); 0: always ``iters`` samples."""
        self.cam = cam
        self.threshold, self.iters, self.seed, self.refine, self.device = threshold, iters, seed, refine, device
        self.prob = prob
        self.mask_match = None
        self.E = None

    def estimatePose(self, kps_ref, kps_cur):
        kps_ref = np.ascontiguousarray(kps_ref, dtype=np.float32).reshape(-1, 2)
        kps_cur = np.ascontiguousarray(kps_cur, dtype=np.float32).reshape(-1, 2)
        n = kps_ref.shape[0]
        assert kps_cur.shape[0] == n
        pts = torch.from_numpy(np.stack([kps_cur, kps_ref])).to(self.device)  # frame 0 = current, 1 = reference
        zero = torch.zeros(1, dtype=torch.int32, device=self.device)
        out = torch_ops.pose_batch(pts, zero, zero + 1, torch.full((1,), n, dtype=torch.int32, device=self.device),
                                   intrinsics=(self.cam.fx, self.cam.fy, self.cam.cx, self.cam.cy),
                                   threshold=self.threshold, iters=self.iters, seed=self.seed, refine=self.refine,
                                   confidence=self.prob)
        self.mask_match = out["mask"][0].cpu().numpy().reshape(-1, 1)
        self.E = out["E"][0].cpu().numpy().astype(np.float64)
        return out["R"][0].cpu().numpy().astype(np.float64), out["t"][0].cpu().numpy().astype(np.float64).reshape(3, 1)
