"""numpy restatement of the relative-pose step (TEST INFRASTRUCTURE; never imported by the product).

The reference calls OpenCV (src/visual_odometry/visual_odometry.py:383-412):

    E, mask = cv2.findEssentialMat(kpn_cur, kpn_ref, focal=1, pp=(0, 0), method=USAC_MSAC|RANSAC, prob=0.999,
                                   threshold=0.0003)
    _, R, t, _ = cv2.recoverPose(E, kpn_cur, kpn_ref, focal=1, pp=(0, 0))

OpenCV (opencv-python 4.10.0.84 in the reference's requirements.txt / uv.lock) is third-party and not part of the
reference tree, so its published algorithms are restated here, deliberately by *different* numerical routes than the
CUDA path so that the two check each other:

  * five-point problem (Nister 2004 / Stewenius 2006): null space by SVD, the ten cubic constraints built with a
    generic polynomial class, Groebner-style elimination and the 10x10 action matrix, solutions from its
    eigenvectors (the CUDA path uses Gauss-Jordan, a trilinear assembly and a tenth-degree polynomial in z);
  * consensus: squared Sampson distance, inlier iff <= threshold^2 (checked against cv2 in tests/test_pose_host.py),
    truncated-cost (MSAC) score;
  * recoverPose: SVD decomposition, four (R, t) candidates, linear (DLT) triangulation of every match and the
    depth test 0 < z < 50 in both cameras, as OpenCV does (the CUDA path intersects the two rays in closed form).

``ransac_pose`` reproduces the product's counter-based sample sequence, so for a given seed it evaluates the same
hypotheses as nvs_pose_batch.  Pinned against OpenCV through tests/golden/pose_cv2.npz (oracle/gen_pose_golden.py).
"""
from __future__ import annotations

import itertools

import numpy as np

M64 = (1 << 64) - 1


# ---- sample sequence (csrc/pose_math.h: mix64 / sample5) -------------------------------------------------------------
def mix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def sample5(seed: int, pair: int, it: int, n: int):
    s = mix64((seed ^ ((pair & 0xFFFFFFFF) << 32) ^ (it & 0xFFFFFFFF)) & M64)
    idx = []
    while len(idx) < 5:
        s = mix64(s)
        c = (s >> 11) % n
        if c not in idx:
            idx.append(c)
    return idx


# ---- polynomials in (x, y, z) as {exponent tuple: coefficient} --------------------------------------------------------
class Poly(dict):
    def __add__(self, o):
        r = Poly(self)
        for k, v in o.items():
            r[k] = r.get(k, 0.0) + v
        return r

    def __sub__(self, o):
        return self + o.scale(-1.0)

    def scale(self, s):
        return Poly({k: v * s for k, v in self.items()})

    def __mul__(self, o):
        r = Poly()
        for (a, va), (b, vb) in itertools.product(self.items(), o.items()):
            k = (a[0] + b[0], a[1] + b[1], a[2] + b[2])
            r[k] = r.get(k, 0.0) + va * vb
        return r


_DEG3 = [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3)]
_BASIS = [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
_MONO = _DEG3 + _BASIS


def five_point(p1: np.ndarray, p2: np.ndarray):
    """p1, p2 (5,2) -> list of 3x3 essential matrices with p2^T E p1 = 0 (Frobenius norm sqrt 2)."""
    p1, p2 = np.asarray(p1, np.float64), np.asarray(p2, np.float64)
    x1, y1, x2, y2 = p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]
    Q = np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, np.ones(5)], 1)
    _, _, Vt = np.linalg.svd(Q)
    B = Vt[5:].reshape(4, 3, 3)  # E = x B0 + y B1 + z B2 + B3
    var = [Poly({(1, 0, 0): 1.0}), Poly({(0, 1, 0): 1.0}), Poly({(0, 0, 1): 1.0}), Poly({(0, 0, 0): 1.0})]
    E = [[Poly() for _ in range(3)] for _ in range(3)]
    for i in range(3):
        for j in range(3):
            for k in range(4):
                E[i][j] = E[i][j] + var[k].scale(B[k, i, j])

    def matmul(A, Bm):
        return [[A[i][0] * Bm[0][j] + A[i][1] * Bm[1][j] + A[i][2] * Bm[2][j] for j in range(3)] for i in range(3)]

    Et = [[E[j][i] for j in range(3)] for i in range(3)]
    EEt = matmul(E, Et)
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    EEtE = matmul(EEt, E)
    eqs = [EEtE[i][j].scale(2.0) - tr * E[i][j] for i in range(3) for j in range(3)]
    det = (E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1]) - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0]) +
           E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]))
    eqs.append(det)
    M = np.array([[q.get(m, 0.0) for m in _MONO] for q in eqs])
    try:
        Bm = np.linalg.solve(M[:, :10], M[:, 10:])  # [I | Bm]
    except np.linalg.LinAlgError:
        return []
    A = np.zeros((10, 10))
    for r, m in enumerate(_BASIS):  # x * basis[r] expressed in the basis
        xm = (m[0] + 1, m[1], m[2])
        if xm in _BASIS:
            A[r, _BASIS.index(xm)] = 1.0
        else:
            A[r] = -Bm[_DEG3.index(xm)]
    w, V = np.linalg.eig(A)
    out = []
    for k in range(10):
        if abs(w[k].imag) > 1e-9 * max(1.0, abs(w[k].real)):
            continue
        v = V[:, k].real
        if abs(v[9]) < 1e-14:
            continue
        x, y, z = v[6] / v[9], v[7] / v[9], v[8] / v[9]
        Ek = x * B[0] + y * B[1] + z * B[2] + B[3]
        out.append(Ek * np.sqrt(2.0) / np.linalg.norm(Ek))
    return out


# ---- consensus -----------------------------------------------------------------------------------------------------
def sampson_sq(E, cur, ref):
    A = np.c_[cur, np.ones(len(cur))].astype(np.float64)
    Bp = np.c_[ref, np.ones(len(ref))].astype(np.float64)
    Ea, Etb = A @ np.asarray(E, np.float64).T, Bp @ np.asarray(E, np.float64)
    r = (Bp * Ea).sum(1)
    return r * r / (Ea[:, 0] ** 2 + Ea[:, 1] ** 2 + Etb[:, 0] ** 2 + Etb[:, 1] ** 2)


def msac_cost(E, cur, ref, thr):
    return float(np.minimum(sampson_sq(E, cur, ref) / thr ** 2, 1.0).sum())


# ---- recoverPose ---------------------------------------------------------------------------------------------------
def decompose_essential(E):
    U, _, Vt = np.linalg.svd(np.asarray(E, np.float64))
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2]


def triangulate_dlt(P0, P1, cur, ref):
    X = np.zeros((len(cur), 4))
    for i, (a, b) in enumerate(zip(cur, ref)):
        A = np.stack([a[0] * P0[2] - P0[0], a[1] * P0[2] - P0[1], b[0] * P1[2] - P1[0], b[1] * P1[2] - P1[1]])
        X[i] = np.linalg.svd(A)[2][-1]
    return X


def recover_pose(E, cur, ref, dist=50.0):
    R1, R2, t = decompose_essential(E)
    cur, ref = np.asarray(cur, np.float64), np.asarray(ref, np.float64)
    P0 = np.eye(3, 4)
    best, best_good = None, -1
    for R, tt in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        P1 = np.c_[R, tt]
        X = triangulate_dlt(P0, P1, cur, ref)
        X = X / X[:, 3:4]
        z0 = X[:, 2]
        z1 = (X @ P1.T)[:, 2]
        good = int(((z0 > 0) & (z0 < dist) & (z1 > 0) & (z1 < dist)).sum())
        if good > best_good:
            best, best_good = (R, tt), good
    return best[0], best[1], best_good


def ransac_pose(cur, ref, thr=0.0003, iters=512, seed=0, pair=0):
    """Same hypotheses as nvs_pose_batch (same sample sequence), scored in fp64; no refinement."""
    cur, ref = np.asarray(cur, np.float32), np.asarray(ref, np.float32)
    n = len(cur)
    if n < 5:
        return None
    best, best_E = np.inf, None
    for it in range(iters):
        idx = sample5(seed, pair, it, n)
        for E in five_point(cur[idx], ref[idx]):
            c = msac_cost(E, cur, ref, thr)
            if c < best:
                best, best_E = c, E
    mask = sampson_sq(best_E, cur, ref) <= thr ** 2
    R, t, _ = recover_pose(best_E, cur, ref)
    return {"E": best_E, "R": R, "t": t, "mask": mask, "inliers": int(mask.sum()), "cost": best}
