"""Golden vectors for the pose step (TEST INFRASTRUCTURE): synthetic two-view scenes and what OpenCV returns for the
reference's call sequence (src/visual_odometry/visual_odometry.py:383-412).

    python oracle/gen_pose_golden.py   ->  tests/golden/pose_cv2.npz

OpenCV (opencv-python 4.10.0.84 in the reference's requirements) is a third-party dependency whose source is not in
the reference tree; the cv2 installed in the build container (4.13.0) has no ``USAC_MSAC`` attribute, so the
reference's own ``try/except`` (:390-393) selects ``cv2.RANSAC`` there -- that is the call recorded here.  RANSAC
draws random samples: agreement with these vectors is by tolerance (tests/test_pose_host.py), not bit-exact.
"""
import os

import cv2
import numpy as np


def scene(n, seed, noise, outlier_frac):
    """Points in front of both cameras, small inter-frame motion (VO-like), normalised image coordinates."""
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(4, 30, n)]
    R, _ = cv2.Rodrigues(rng.normal(0, 0.02, 3))
    t = np.array([rng.normal(0, 0.1), rng.normal(0, 0.05), 1.0])
    t /= np.linalg.norm(t)
    x1 = X[:, :2] / X[:, 2:3]
    X2 = (R @ X.T).T + t
    x2 = X2[:, :2] / X2[:, 2:3]
    x1 = x1 + rng.normal(0, noise, x1.shape)
    x2 = x2 + rng.normal(0, noise, x2.shape)
    out = rng.random(n) < outlier_frac
    x2[out] = rng.uniform(-0.5, 0.5, (int(out.sum()), 2))
    return x1.astype(np.float32), x2.astype(np.float32), R, t, out


def main():
    cases = [(1500, 0, 0.0, 0.3), (1500, 1, 0.0, 0.5), (400, 2, 0.0, 0.2), (1500, 3, 0.3 / 700, 0.25),
             (4000, 4, 0.3 / 700, 0.25), (300, 5, 0.2 / 700, 0.4), (60, 6, 0.0, 0.1), (1500, 7, 0.5 / 700, 0.1)]
    z = {"n_cases": np.int64(len(cases))}
    cv2.setRNGSeed(12345)
    for i, (n, seed, noise, frac) in enumerate(cases):
        cur, ref, R, t, out = scene(n, seed, noise, frac)
        E, mask = cv2.findEssentialMat(cur, ref, focal=1, pp=(0.0, 0.0), method=cv2.RANSAC, prob=0.999,
                                       threshold=0.0003)
        _, Rc, tc, _ = cv2.recoverPose(E[:3], cur, ref, focal=1, pp=(0.0, 0.0))
        z.update({f"cur{i}": cur, f"ref{i}": ref, f"R_true{i}": R, f"t_true{i}": t, f"outlier{i}": out,
                  f"noise{i}": np.float64(noise), f"E_cv{i}": E[:3], f"mask_cv{i}": mask.ravel().astype(np.uint8),
                  f"R_cv{i}": Rc, f"t_cv{i}": tc.ravel()})
        print(i, n, "cv2 inliers", int(mask.sum()), "true inliers", int((~out).sum()))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pose_cv2.npz")
    np.savez_compressed(path, **z)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
