"""BfFeatureMatcher -- device version of src/visual_odometry/feature_matcher.py:234-248 (+ :89-98, :179-209).

``match(des1, des2)`` accepts numpy arrays (as the reference does) or CUDA tensors and returns the
reference's ``(idx1, idx2, score)`` python lists.  ``match_device`` returns device tensors.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

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
        return ops.match(self._dev(des1), self._dev(des2), ratio=ratio, mode=mode)

    def knn_match(self, des1, des2):
        """cv2.BFMatcher.knnMatch(des1, des2, k=2) as (idx (n1,2), dist (n1,2)) device tensors."""
        return ops.match(self._dev(des1), self._dev(des2), mode=2)

    def match(self, des1, des2, ratio_test=None):
        i1, i2, dd, cnt = self.match_device(des1, des2, ratio_test)
        n = int(cnt)
        return i1[:n].cpu().tolist(), i2[:n].cpu().tolist(), dd[:n].cpu().tolist()
