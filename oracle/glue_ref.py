"""CPU restatements of the numeric glue either side of the network (TEST INFRASTRUCTURE).

* ``frontend_decode``   -- src/visual_odometry/frontend.py:94-126 (threshold + argpartition top-k)
* ``select_k_best``     -- src/evaluation/descriptor.py:12-36
* ``knn2_l2`` + ``good_matches_one_to_one`` -- src/visual_odometry/feature_matcher.py:89-98,179-209
  (the 2-NN itself is ``cv2.BFMatcher(NORM_L2).knnMatch(k=2)``: OpenCV 4.10.0.84 per requirements.txt,
  third-party; restated as exact brute force, validated against cv2 4.13 in tests)
* ``mutual_nn``         -- ``cv2.BFMatcher(NORM_L2, crossCheck=True).match`` (evaluation/descriptor.py:221)
* ``flat_l2_search``    -- ``faiss.IndexFlatL2.add/search`` (src/evaluation/global_descriptor.py:55-60;
  faiss-cpu 1.9.0 per requirements.txt, third-party, NOT installable here -> parity unpinned).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Optional, Sequence, Tuple

import numpy as np
import torch


def frontend_decode(post: dict, nfeatures: int, thresh: float = 0.7, top_k: int = 4000,
                    classes_to_filter: Optional[Sequence[int]] = None):
    """frontend.py:94-126 for ONE frame (the reference glue assumes B == 1).

    Returns (pts (n,2) xy, desc (n,nfeatures), seg labels, kept cell indices).  The kept set is what
    parity is judged on: ``argpartition`` returns it unordered and with arbitrary tie choice.
    """
    score, coord, feat, seg = post["score"], post["coord"], post["feat"], post["seg"]
    assert score.shape[0] == 1
    sc = torch.cat([coord, score], dim=1).view(3, -1).t().cpu().numpy()  # :94
    ft = feat.view(nfeatures, -1).t().cpu().numpy()  # :106
    mask = sc[:, 2] > thresh  # :108
    segv = seg.view(-1).cpu().numpy()
    if classes_to_filter is not None:  # :109-114 (requires sample_segmentation so seg is per cell)
        mask = mask & ~np.isin(segv, list(classes_to_filter))
        segv = segv[mask]
    cells = np.nonzero(mask)[0]
    ft, pts, s = ft[mask, :], sc[mask, :2], sc[mask, 2]
    if len(s) > top_k and top_k > 0:  # :119-126
        sel = np.argpartition(s, -top_k)[-top_k:]
        pts, ft, cells = pts[sel], ft[sel], cells[sel]
        if classes_to_filter is not None:
            segv = segv[sel]
    return pts.copy(), ft.copy(), segv, cells


def select_k_best(points: np.ndarray, descriptors: np.ndarray, k: int):
    """evaluation/descriptor.py:31-35: ascending argsort by probability, keep the last k."""
    order = points[:, 2].argsort()
    start = min(k, points.shape[0])
    return points[order, :2][-start:], descriptors[order][-start:]


def knn2_l2(des1: np.ndarray, des2: np.ndarray):
    """Exact 2-nearest neighbours under L2 (NOT squared) = knnMatch(des1, des2, k=2) with NORM_L2.

    Returns (idx (n1,2) int64, dist (n1,2) float32); ties -> lowest train index first.
    """
    a = torch.from_numpy(np.ascontiguousarray(des1)).float()
    b = torch.from_numpy(np.ascontiguousarray(des2)).float()
    d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1) if a.shape[0] * b.shape[0] <= 4_000_000 else None
    if d2 is None:
        d2 = torch.cdist(a.double(), b.double()).pow(2).float()
    d = d2.sqrt()
    dist, idx = torch.topk(d, k=min(2, b.shape[0]), dim=1, largest=False, sorted=True)
    return idx.numpy().astype(np.int64), dist.numpy().astype(np.float32)


def good_matches_one_to_one(idx: np.ndarray, dist: np.ndarray, ratio_test: float = 0.7):
    """feature_matcher.py:179-209 restated over the (idx, dist) arrays of the 2-NN search."""
    float_inf = float("inf")
    dist_match = defaultdict(lambda: float_inf)
    index_match = {}
    idx1, idx2, score = [], [], []
    for q in range(idx.shape[0]):
        m_d, n_d, t = float(dist[q, 0]), float(dist[q, 1]), int(idx[q, 0])
        if m_d > ratio_test * n_d:  # :193
            continue
        cur = dist_match[t]
        if cur == float_inf:  # first claim of this train index
            dist_match[t] = m_d
            idx1.append(q)
            idx2.append(t)
            index_match[t] = len(idx2) - 1
            score.append(m_d)
        elif m_d < cur:
            # NB the reference does not update dist_match here (:201-208): later queries are compared
            # with the FIRST stored distance, not the best so far.  Restated literally.
            i = index_match[t]
            idx1[i], idx2[i], score[i] = q, t, m_d
    return idx1, idx2, score


def bf_match(des1: np.ndarray, des2: np.ndarray, ratio_test: float = 0.7):
    """BfFeatureMatcher.match (feature_matcher.py:89-98)."""
    idx, dist = knn2_l2(des1, des2)
    return good_matches_one_to_one(idx, dist, ratio_test)


def mutual_nn(des1: np.ndarray, des2: np.ndarray):
    """cv2.BFMatcher(NORM_L2, crossCheck=True).match: (i, j) kept iff j = nn(i) and i = nn(j)."""
    a = torch.from_numpy(np.ascontiguousarray(des1)).float()
    b = torch.from_numpy(np.ascontiguousarray(des2)).float()
    d = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
    nn12 = d.argmin(dim=1)
    nn21 = d.argmin(dim=0)
    i = torch.arange(a.shape[0])
    keep = nn21[nn12] == i
    return i[keep].numpy(), nn12[keep].numpy(), d[i[keep], nn12[keep]].sqrt().numpy()


def flat_l2_search(db: torch.Tensor, q: torch.Tensor, k: int, block: int = 256,
                   id_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """IndexFlatL2.search: exact squared L2 ``|q|^2 + |x|^2 - 2 q.x`` in fp32, k smallest ascending,
    int64 labels (global_descriptor.py:55-60 call site; faiss semantics)."""
    db = db.float()
    q = q.float()
    xn = (db * db).sum(1)
    D = torch.empty(q.shape[0], k, dtype=torch.float32)
    I = torch.empty(q.shape[0], k, dtype=torch.int64)
    for s in range(0, q.shape[0], block):
        qq = q[s:s + block]
        d2 = (qq * qq).sum(1, keepdim=True) + xn.unsqueeze(0) - 2.0 * (qq @ db.t())
        dv, iv = torch.topk(d2, k, dim=1, largest=False, sorted=True)
        D[s:s + block], I[s:s + block] = dv, iv + id_offset
    return D, I


def merge_shard_topk(D_parts: Sequence[torch.Tensor], I_parts: Sequence[torch.Tensor], k: int):
    """Merge per-shard (Q,k) results (global ids) into the global top-k, ascending distance."""
    D = torch.cat(list(D_parts), dim=1)
    I = torch.cat(list(I_parts), dim=1)
    dv, pos = torch.topk(D, k, dim=1, largest=False, sorted=True)
    return dv, torch.gather(I, 1, pos)


def preprocess_u8(img_hwc_u8: np.ndarray, new_size=None):
    """Input side of the VO loop as the reference does it on the host (visual_odometry.py:281-291, frontend.py:79):
    kornia.image_to_tensor(image).float() / 255 -> kornia.geometry.transform.resize(size=new_size) (bilinear,
    align_corners=None -> False, no antialias: a thin wrapper over F.interpolate; kornia is not installed here,
    so this line of the restatement is UNPINNED) -> sub(0.5).mul(2).  Returns (3,H',W') float32."""
    import torch
    import torch.nn.functional as F

    t = torch.from_numpy(np.ascontiguousarray(img_hwc_u8)).permute(2, 0, 1).float() / 255.0
    if new_size is not None:
        t = F.interpolate(t.unsqueeze(0), size=tuple(new_size), mode="bilinear", align_corners=False)[0]
    return t.sub(0.5).mul(2.0).numpy()
