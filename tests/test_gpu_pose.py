"""Batched pose kernels (csrc/pose.cu) vs the host harness of the same algebra (same seed -> same hypotheses) and vs
OpenCV's recorded results (tests/golden/pose_cv2.npz); visual_odometry.py:383-412."""
import os

import numpy as np
import pytest
import torch

from pose_util import dir_angle_deg, host_pose, rot_angle_deg, sampson_sq

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose_cv2.npz")


def _batch(z, idxs):
    """Scenes of different lengths packed as 2P 'frames' (frame p = current, P+p = reference) of kmax rows."""
    P = len(idxs)
    kmax = max(len(z[f"cur{i}"]) for i in idxs)
    pts = np.zeros((2 * P, kmax, 2), np.float32)
    cnt = np.zeros(P, np.int32)
    for p, i in enumerate(idxs):
        n = len(z[f"cur{i}"])
        pts[p, :n], pts[P + p, :n], cnt[p] = z[f"cur{i}"], z[f"ref{i}"], n
    return torch.from_numpy(pts).cuda(), torch.from_numpy(cnt).cuda(), kmax


def test_pose_batch_matches_host_harness_and_opencv():
    from nano_vs_slam_b200 import ops

    z = np.load(GOLD)
    idxs = list(range(int(z["n_cases"])))
    P = len(idxs)
    pts, cnt, kmax = _batch(z, idxs)
    a = torch.arange(P, dtype=torch.int32, device="cuda")
    out = ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=512, seed=7)
    again = ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=512, seed=7)
    for k in out:
        assert torch.equal(out[k], again[k]), k                   # deterministic per seed
    E, R, t = out["E"].cpu().numpy(), out["R"].cpu().numpy(), out["t"].cpu().numpy()
    mask, inl = out["mask"].cpu().numpy(), out["inliers"].cpu().numpy()
    for p, i in enumerate(idxs):
        cur, ref = z[f"cur{i}"], z[f"ref{i}"]
        n = len(cur)
        h = host_pose(cur, ref, seed=7, pair=p)
        assert inl[p] == mask[p, :n].sum() and not mask[p, n:].any()
        # same samples, same algebra; fp contraction differs between nvcc and g++, so compare by tolerance
        assert (mask[p, :n] != h["mask"]).mean() < 5e-3, (i, (mask[p, :n] != h["mask"]).sum())
        s = 1.0 if np.abs(E[p] - h["E"]).max() < np.abs(E[p] + h["E"]).max() else -1.0
        assert np.abs(E[p] - s * h["E"]).max() < 1e-4, i
        assert rot_angle_deg(R[p], h["R"]) < 2e-3 and dir_angle_deg(t[p], h["t"]) < 1e-2, i
        assert np.array_equal(mask[p, :n].astype(bool), sampson_sq(E[p], cur, ref) <= 0.0003 ** 2) or \
            (mask[p, :n].astype(bool) != (sampson_sq(E[p], cur, ref) <= 0.0003 ** 2)).mean() < 2e-3
        dR, dt = rot_angle_deg(R[p], z[f"R_cv{i}"]), dir_angle_deg(t[p], z[f"t_cv{i}"])
        if float(z[f"noise{i}"]) == 0.0:
            assert np.array_equal(mask[p, :n], z[f"mask_cv{i}"]), i
            assert dR < 2e-3 and dt < 0.01, (i, dR, dt)
        else:
            assert dR < 0.2 and dt < 3.0, (i, dR, dt)
            assert 0.85 < inl[p] / z[f"mask_cv{i}"].sum() < 1.15
        if i in (2, 6):
            # the independent numpy oracle (action-matrix five-point, DLT cheirality) at the same seed
            from oracle import pose_ref as O

            o = O.ransac_pose(cur, ref, seed=7, pair=p)
            assert (o["mask"] != mask[p, :n].astype(bool)).mean() < 5e-3 and abs(o["inliers"] - int(inl[p])) <= 1
            assert rot_angle_deg(R[p], o["R"]) < 5e-3 and dir_angle_deg(t[p], o["t"]) < 1e-2, i


def test_pose_pixel_coordinates_indices_and_degenerate_pairs():
    from nano_vs_slam_b200 import ops

    z = np.load(GOLD)
    cur, ref = z["cur0"], z["ref0"]
    n = len(cur)
    fx, fy, cx, cy = 707.09, 705.5, 601.9, 183.1
    rng = np.random.default_rng(0)
    perm1, perm2 = rng.permutation(n), rng.permutation(n)
    kmax = n + 37
    pts = np.zeros((3, kmax, 2), np.float32)
    pts[0, perm1] = cur * [fx, fy] + [cx, cy]          # keypoint lists in arbitrary order, pixel coordinates
    pts[1, perm2] = ref * [fx, fy] + [cx, cy]
    idx1 = np.zeros((4, kmax), np.int32)
    idx2 = np.zeros((4, kmax), np.int32)
    idx1[0, :n], idx2[0, :n] = perm1, perm2
    idx1[1, :n], idx2[1, :n] = perm1, perm2
    count = np.array([n, 4, 0, 0], np.int32)            # pair 1: < 5 matches, pairs 2-3: none
    out = ops.pose_batch(torch.from_numpy(pts).cuda(), torch.tensor([0, 0, 2, 0], dtype=torch.int32),
                         torch.tensor([1, 1, 2, 1], dtype=torch.int32), torch.from_numpy(count),
                         torch.from_numpy(idx1).cuda(), torch.from_numpy(idx2).cuda(), intrinsics=(fx, fy, cx, cy), seed=3)
    R, t, inl = out["R"].cpu().numpy(), out["t"].cpu().numpy(), out["inliers"].cpu().numpy()
    assert rot_angle_deg(R[0], z["R_true0"]) < 0.02 and dir_angle_deg(t[0], z["t_true0"]) < 0.05
    assert abs(int(inl[0]) - int((~z["outlier0"]).sum())) <= 3
    for p in (1, 2, 3):
        assert inl[p] == 0 and np.array_equal(R[p], np.eye(3, dtype=np.float32)) and not t[p].any()
        assert not out["mask"][p].any() and not out["E"][p].any()


def test_pose_estimator_and_network_chain():
    """Reference-shaped estimatePose(kps_ref, kps_cur) and the select -> match -> pose chain on network output."""
    import contextlib
    import io
    from types import SimpleNamespace

    from nano_vs_slam_b200 import tiny_factory
    from nano_vs_slam_b200.frontend import KP2DtinyFrontend
    from nano_vs_slam_b200.matcher import PoseEstimator, pose_consecutive
    from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames

    z = np.load(GOLD)
    cam = SimpleNamespace(fx=700.0, fy=700.0, cx=320.0, cy=120.0)
    est = PoseEstimator(cam, seed=11)
    R, t = est.estimatePose(z["ref1"] * 700.0 + [320.0, 120.0], z["cur1"] * 700.0 + [320.0, 120.0])
    assert R.shape == (3, 3) and t.shape == (3, 1) and est.mask_match.shape == (len(z["cur1"]), 1)
    assert rot_angle_deg(R, z["R_true1"]) < 0.02 and dir_angle_deg(t, z["t_true1"]) < 0.05
    assert abs(int(est.mask_match.sum()) - int((~z["outlier1"]).sum())) <= 3

    with contextlib.redirect_stdout(io.StringIO()):
        sd = spread_init(tiny_factory("S", 19).state_dict(), 4321)
        fe = KP2DtinyFrontend(config="S", nClasses=19, nn_thresh=0.0, top_k=300, device="cuda", state_dict=sd)
    x = synthetic_frames(1, 120, 160, 5)
    frames = torch.cat([torch.roll(x, shifts=(0, 2 * i), dims=(2, 3)) for i in range(4)])
    sel, _ = fe.run_batch(frames.cuda(), normalized=True)
    (i1, i2, dd, cnt), pose = pose_consecutive(sel, (100.0, 100.0, 80.0, 60.0), cross_check=True, threshold=0.003)
    assert pose["R"].shape == (3, 3, 3) and int(cnt.min()) > 20
    # the same gathered matches through the host harness give the same consensus
    for p in range(3):
        m = int(cnt[p])
        cur = sel["pts"][p + 1][i1[p, :m].long()].cpu().numpy()
        ref = sel["pts"][p][i2[p, :m].long()].cpu().numpy()
        cur = ((cur - [80.0, 60.0]) / 100.0).astype(np.float32)
        ref = ((ref - [80.0, 60.0]) / 100.0).astype(np.float32)
        h = host_pose(cur, ref, thr=0.003, iters=512, seed=0, pair=p)
        assert abs(int(pose["inliers"][p]) - h["inliers"]) <= max(2, h["inliers"] // 50)


def test_pose_refinement_matches_host_harness():
    from nano_vs_slam_b200 import ops

    z = np.load(GOLD)
    idxs = list(range(int(z["n_cases"])))
    P = len(idxs)
    pts, cnt, kmax = _batch(z, idxs)
    a = torch.arange(P, dtype=torch.int32, device="cuda")
    base = ops.pose_batch(pts, a, a + P, cnt, iters=512, seed=7)
    out = ops.pose_batch(pts, a, a + P, cnt, iters=512, seed=7, refine=10)
    R, t, inl = out["R"].cpu().numpy(), out["t"].cpu().numpy(), out["inliers"].cpu().numpy()
    E, mask = out["E"].cpu().numpy(), out["mask"].cpu().numpy()
    thr2 = 0.0003 ** 2
    for p, i in enumerate(idxs):
        cur, ref = z[f"cur{i}"], z[f"ref{i}"]
        n = len(cur)
        h = host_pose(cur, ref, seed=7, pair=p, refine=10)
        assert rot_angle_deg(R[p], h["R"]) < 2e-3 and dir_angle_deg(t[p], h["t"]) < 2e-2, i
        assert abs(int(inl[p]) - h["inliers"]) <= max(2, h["inliers"] // 100), (i, inl[p], h["inliers"])
        assert inl[p] == mask[p, :n].sum() and not mask[p, n:].any()
        c0 = np.minimum(sampson_sq(base["E"][p].cpu().numpy(), cur, ref), thr2).sum()
        c1 = np.minimum(sampson_sq(E[p], cur, ref), thr2).sum()
        assert c1 <= c0 * (1 + 1e-5), (i, c0, c1)


def test_pose_adaptive_sample_count():
    """confidence = 0.999 (the reference's prob): clean scenes stop after the first round(s), a pair that stops after m
    samples returns exactly what the fixed-count call returns with iters = m (same sample sequence), and the estimate
    stays within the golden tolerances."""
    from nano_vs_slam_b200 import ops

    z = np.load(GOLD)
    idxs = list(range(int(z["n_cases"])))
    P = len(idxs)
    pts, cnt, kmax = _batch(z, idxs)
    a = torch.arange(P, dtype=torch.int32, device="cuda")
    out = ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=512, seed=7, confidence=0.999, round_size=32)
    used = out["iters"].cpu().numpy()
    assert (used % 32 == 0).all() and (used >= 32).all() and (used <= 512).all()
    assert used.min() < 512  # at least the noise-free scenes (>= 70 % inliers) finish early
    again = ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=512, seed=7, confidence=0.999, round_size=32)
    for k in out:
        assert torch.equal(out[k], again[k]), k
    R, t, inl = out["R"].cpu().numpy(), out["t"].cpu().numpy(), out["inliers"].cpu().numpy()
    for p, i in enumerate(idxs):
        fixed = ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=int(used[p]), seed=7)
        for k in ("E", "R", "t", "mask", "inliers"):
            assert torch.equal(out[k][p], fixed[k][p]), (i, k)
        dR, dt = rot_angle_deg(R[p], z[f"R_cv{i}"]), dir_angle_deg(t[p], z[f"t_cv{i}"])
        if float(z[f"noise{i}"]) == 0.0:
            assert dR < 2e-3 and dt < 0.01, (i, dR, dt, used[p])
        else:
            assert dR < 0.3 and dt < 4.0, (i, dR, dt, used[p])
            assert 0.8 < inl[p] / z[f"mask_cv{i}"].sum() < 1.2
    # through the custom operator
    from nano_vs_slam_b200 import torch_ops

    o2 = torch_ops.pose_batch(pts, a, a + P, cnt, threshold=0.0003, iters=512, seed=7, confidence=0.999, round_size=32)
    assert torch.equal(o2["iters"], out["iters"]) and torch.equal(o2["R"], out["R"])
