"""Pose step, checked on the CPU: csrc/pose_math.h (the algebra the pose kernels run) is compiled with g++ through the
host harness tests/host/pose_host.cpp and compared with OpenCV's results for the reference's call sequence
(visual_odometry.py:383-412), recorded in tests/golden/pose_cv2.npz by oracle/gen_pose_golden.py.

RANSAC draws random samples, so agreement with cv2 is by tolerance: on noise-free matches with gross outliers both
find the exact epipolar geometry (identical inlier sets); with pixel noise above the 0.0003 threshold both return a
minimal-sample estimate and are compared through their distance to each other and to the ground truth."""
import os

import numpy as np

from pose_util import dir_angle_deg, five_point, host_pose, real_roots, rot_angle_deg, sampson_sq

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose_cv2.npz")


def _scene(n, seed, noise=0.0):
    rng = np.random.default_rng(seed)
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(4, 30, n)]
    w = rng.normal(0, 0.05, 3)
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    t = rng.normal(0, 1, 3)
    t /= np.linalg.norm(t)
    x1 = X[:, :2] / X[:, 2:3]
    X2 = (R @ X.T).T + t
    x2 = X2[:, :2] / X2[:, 2:3]
    return x1 + rng.normal(0, noise, x1.shape), x2 + rng.normal(0, noise, x2.shape), R, t


def test_real_roots_match_numpy():
    rng = np.random.default_rng(0)
    for trial in range(200):
        deg = int(rng.integers(1, 11))
        roots_true = rng.normal(0, 2, deg)
        if trial % 3 == 0 and deg >= 4:  # complex pairs: fewer real roots
            p = np.poly(roots_true[: deg - 2])
            p = np.polymul(p, [1.0, rng.normal(), abs(rng.normal()) + 2.0])
            roots_true = roots_true[: deg - 2]
        else:
            p = np.poly(roots_true)
        got = np.sort(real_roots(p[::-1] * rng.uniform(0.5, 2.0)))
        ref = np.sort(roots_true)
        if np.min(np.diff(ref)) < 1e-3 if len(ref) > 1 else False:
            continue  # near-double roots: bracketing by sign change is not expected to split them
        assert len(got) == len(ref), (trial, got, ref)
        assert np.allclose(got, ref, atol=1e-7, rtol=1e-7), (trial, got, ref)


def test_five_point_candidates_are_essential_and_contain_the_truth():
    for seed in range(40):
        x1, x2, R, t, = _scene(5, seed)
        Es = five_point(x1, x2)
        assert 1 <= len(Es) <= 10
        Et = np.cross(np.eye(3), t) @ R
        Et *= np.sqrt(2) / np.linalg.norm(Et)
        best = 1e9
        for E in Es:
            assert abs(np.linalg.det(E)) < 1e-6
            assert np.abs(2 * E @ E.T @ E - np.trace(E @ E.T) * E).max() < 1e-6
            assert max(abs(np.r_[x2[k], 1] @ E @ np.r_[x1[k], 1]) for k in range(5)) < 1e-10
            best = min(best, np.abs(E - Et).max(), np.abs(E + Et).max())
        assert best < 1e-6, (seed, best)


def test_pose_matches_opencv_golden():
    z = np.load(GOLD)
    for i in range(int(z["n_cases"])):
        cur, ref = z[f"cur{i}"], z[f"ref{i}"]
        o = host_pose(cur, ref, seed=7, pair=i)
        mc = z[f"mask_cv{i}"].astype(bool)
        mo = o["mask"].astype(bool)
        assert abs(np.linalg.det(o["R"].astype(np.float64)) - 1) < 1e-5 and abs(np.linalg.norm(o["t"]) - 1) < 1e-5
        # the returned mask is the thresholded Sampson distance of the returned E (cv2 convention: <= thr^2)
        assert np.array_equal(mo, sampson_sq(o["E"], cur, ref) <= np.float32(0.0003) ** 2) or \
            (mo != (sampson_sq(o["E"], cur, ref) <= 0.0003 ** 2)).mean() < 2e-3
        dR, dt = rot_angle_deg(o["R"], z[f"R_cv{i}"]), dir_angle_deg(o["t"], z[f"t_cv{i}"])
        if float(z[f"noise{i}"]) == 0.0:
            assert np.array_equal(mo, mc), i                     # identical consensus set
            assert dR < 2e-3 and dt < 0.01, (i, dR, dt)
        else:
            # minimal-sample estimates under noise above the threshold: same order of accuracy as OpenCV's
            assert dR < 0.2 and dt < 3.0, (i, dR, dt)
            assert 0.85 < mo.sum() / mc.sum() < 1.15, (i, mo.sum(), mc.sum())
            eR_o, eR_c = rot_angle_deg(o["R"], z[f"R_true{i}"]), rot_angle_deg(z[f"R_cv{i}"], z[f"R_true{i}"])
            et_o, et_c = dir_angle_deg(o["t"], z[f"t_true{i}"]), dir_angle_deg(z[f"t_cv{i}"], z[f"t_true{i}"])
            assert eR_o < 3 * eR_c + 0.05 and et_o < 3 * et_c + 0.5, (i, eR_o, eR_c, et_o, et_c)


def test_pose_is_deterministic_per_seed_and_handles_general_motion():
    x1, x2, R, t = _scene(600, 3, noise=1e-5)
    a = host_pose(x1, x2, seed=1)
    b = host_pose(x1, x2, seed=1)
    c = host_pose(x1, x2, seed=2)
    assert np.array_equal(a["E"], b["E"]) and np.array_equal(a["mask"], b["mask"])
    for o in (a, c):
        assert rot_angle_deg(o["R"], R) < 0.02 and dir_angle_deg(o["t"], t) < 0.05
        assert o["inliers"] == 600


def test_pose_too_few_matches():
    x1, x2, _, _ = _scene(4, 0)
    assert host_pose(x1, x2)["inliers"] == -1


def test_refinement_lowers_truncated_cost_and_keeps_the_geometry():
    """refine > 0: Gauss-Newton on the consensus set (the LO / polishing role of USAC_MSAC).  The truncated Sampson
    cost of the returned E never increases, more matches become inliers under noise, the pose stays near the truth."""
    z = np.load(GOLD)
    thr2 = 0.0003 ** 2
    for i in range(int(z["n_cases"])):
        cur, ref = z[f"cur{i}"], z[f"ref{i}"]
        a = host_pose(cur, ref, seed=7, pair=i)
        b = host_pose(cur, ref, seed=7, pair=i, refine=10)
        ca = np.minimum(sampson_sq(a["E"], cur, ref), thr2).sum()
        cb = np.minimum(sampson_sq(b["E"], cur, ref), thr2).sum()
        assert cb <= ca * (1 + 1e-6), (i, ca, cb)
        assert abs(np.linalg.det(b["R"].astype(np.float64)) - 1) < 1e-5 and abs(np.linalg.norm(b["t"]) - 1) < 1e-5
        assert b["inliers"] >= a["inliers"] - 1, (i, a["inliers"], b["inliers"])
        assert rot_angle_deg(b["R"], z[f"R_true{i}"]) < 0.1 and dir_angle_deg(b["t"], z[f"t_true{i}"]) < 1.5, i
        if float(z[f"noise{i}"]) == 0.0:
            assert rot_angle_deg(b["R"], z[f"R_cv{i}"]) < 0.02 and dir_angle_deg(b["t"], z[f"t_cv{i}"]) < 0.02


# ---- the numpy oracle (oracle/pose_ref.py): pinned to OpenCV, then used to check the product algebra -----------------
def test_oracle_consensus_rule_and_recover_pose_equal_opencv():
    """Given OpenCV's own E: its mask is exactly `squared Sampson distance <= threshold^2`, and the oracle's
    recoverPose restatement (SVD decomposition, DLT triangulation, depth test) returns OpenCV's R and t."""
    from oracle import pose_ref as O

    z = np.load(GOLD)
    for i in range(int(z["n_cases"])):
        cur, ref, E = z[f"cur{i}"], z[f"ref{i}"], z[f"E_cv{i}"]
        err = O.sampson_sq(E, cur, ref)
        mc = z[f"mask_cv{i}"].astype(bool)
        border = np.abs(err - 0.0003 ** 2) < 1e-12          # cv2 evaluates in float; skip exact-threshold ties
        assert np.array_equal((err <= 0.0003 ** 2)[~border], mc[~border]), i
        if len(cur) <= 1500:
            R, t, good = O.recover_pose(E, cur, ref)
            assert rot_angle_deg(R, z[f"R_cv{i}"]) < 1e-5 and dir_angle_deg(t, z[f"t_cv{i}"]) < 1e-5, i


def test_five_point_equals_oracle_action_matrix_solver():
    """Two independent routes to the same solution set: tenth-degree polynomial (product) vs action-matrix eigenvectors."""
    from oracle import pose_ref as O

    def canon(E):
        E = E / np.linalg.norm(E)
        return E * np.sign(E.flat[np.argmax(np.abs(E))])

    n_checked = 0
    for seed in range(30):
        x1, x2, _, _ = _scene(5, 100 + seed)
        a = sorted((canon(E) for E in five_point(x1, x2)), key=lambda m: tuple(np.round(m.ravel(), 5)))
        b = sorted((canon(E) for E in O.five_point(x1, x2)), key=lambda m: tuple(np.round(m.ravel(), 5)))
        if len(a) != len(b):
            continue  # a near-double root may be split by one solver and merged by the other
        n_checked += 1
        assert max(np.abs(p - q).max() for p, q in zip(a, b)) < 1e-6, seed
    assert n_checked >= 25


def test_same_seed_pipeline_equals_oracle():
    from oracle import pose_ref as O

    z = np.load(GOLD)
    for i in (2, 5, 6):
        cur, ref = z[f"cur{i}"], z[f"ref{i}"]
        o = O.ransac_pose(cur, ref, seed=7, pair=i)
        h = host_pose(cur, ref, seed=7, pair=i)
        assert (o["mask"] != h["mask"].astype(bool)).mean() < 5e-3 and abs(o["inliers"] - h["inliers"]) <= 1
        assert rot_angle_deg(o["R"], h["R"]) < 1e-3 and dir_angle_deg(o["t"], h["t"]) < 5e-3, i


def test_planar_scene_and_large_rotation():
    """A dominant plane (road / facade) plus some structure off it, large inter-frame rotation.  (A purely planar scene
    has two essential matrices that explain every match -- the twisted-pair of the plane homography -- so only scenes
    with some depth relief determine the pose; minimal samples drawn entirely from the plane still produce the correct
    E among their candidates, which the consensus over the off-plane matches then selects.)"""
    rng = np.random.default_rng(11)
    n = 500
    # all points on the plane z = 8 + 0.3 x (a road / facade), sideways + forward motion, 12 degree rotation
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), np.zeros(n)]
    X[:, 2] = 8.0 + 0.3 * X[:, 0]
    off = rng.random(n) < 0.2
    X[off, 2] += rng.uniform(-3, 6, int(off.sum()))
    w = np.array([0.05, 0.2, -0.03])
    th = np.linalg.norm(w)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    t = np.array([0.8, 0.1, 0.59])
    t /= np.linalg.norm(t)
    x1 = X[:, :2] / X[:, 2:3]
    X2 = (R @ X.T).T + t
    x2 = X2[:, :2] / X2[:, 2:3]
    out = rng.random(n) < 0.3
    x2[out] = rng.uniform(-0.6, 0.6, (int(out.sum()), 2))
    for refine in (0, 10):
        o = host_pose(x1, x2, seed=5, refine=refine)
        assert abs(o["inliers"] - int((~out).sum())) <= 3
        assert rot_angle_deg(o["R"], R) < 0.01 and dir_angle_deg(o["t"], t) < 0.05, refine


def test_five_point_degenerate_samples_stay_finite():
    """Duplicated / collinear / coincident / tiny minimal samples: never a hang, never a non-finite candidate, every
    returned E satisfies the five epipolar constraints; stationary frames give 'no model' (kernel: R = I, t = 0)."""
    rng = np.random.default_rng(123)
    for trial in range(1200):
        kind = trial % 6
        p1, p2 = rng.uniform(-1, 1, (5, 2)), rng.uniform(-1, 1, (5, 2))
        if kind == 1:
            p2 = p1.copy()
        elif kind == 2:
            p1[1], p2[1] = p1[0], p2[0]
        elif kind == 3:
            p1[:, 1], p2[:, 1] = 0.3 * p1[:, 0], 0.3 * p2[:, 0]
        elif kind == 4:
            p1, p2 = p1 * 1e-4, p2 * 1e-4
        elif kind == 5:
            p2 = p1 + 1e-9 * rng.normal(size=(5, 2))
        Es = five_point(p1, p2)
        assert len(Es) <= 10 and np.isfinite(Es).all()
        for E in Es:
            assert max(abs(np.r_[p2[k], 1] @ E @ np.r_[p1[k], 1]) for k in range(5)) < 1e-6, (trial, kind)
    a = rng.uniform(-1, 1, (200, 2)).astype(np.float32)
    assert host_pose(a, a.copy(), seed=1)["inliers"] == -2
