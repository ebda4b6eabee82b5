"""Helpers shared by the pose tests: builds tests/host/pose_host.cpp (host harness of csrc/pose_math.h) with g++."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def host_lib():
    global _lib
    if _lib is None:
        out = os.path.join(tempfile.gettempdir(), f"nvs_pose_host_{os.getuid()}_{os.getpid()}.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host", "pose_host.cpp")])
        _lib = ctypes.CDLL(out)
        os.unlink(out)
        _lib.nvs_host_pose.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                       ctypes.c_uint64, ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int]
        _lib.nvs_host_five_point.argtypes = [ctypes.c_void_p] * 3
        _lib.nvs_host_real_roots.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    return _lib


def host_pose(cur, ref, thr=0.0003, iters=512, seed=0, pair=0, refine=0):
    cur = np.ascontiguousarray(cur, np.float32)
    ref = np.ascontiguousarray(ref, np.float32)
    n = len(cur)
    E, R, t, mask = np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(3, np.float32), np.zeros(n, np.uint8)
    ninl = host_lib().nvs_host_pose(cur.ctypes.data, ref.ctypes.data, n, thr, iters, seed, pair, E.ctypes.data,
                                    R.ctypes.data, t.ctypes.data, mask.ctypes.data, refine)
    return {"E": E.reshape(3, 3), "R": R.reshape(3, 3), "t": t, "mask": mask, "inliers": ninl}


def five_point(p1, p2):
    p1 = np.ascontiguousarray(p1, np.float64)
    p2 = np.ascontiguousarray(p2, np.float64)
    Es = np.zeros((10, 9))
    n = host_lib().nvs_host_five_point(p1.ctypes.data, p2.ctypes.data, Es.ctypes.data)
    return Es[:n].reshape(n, 3, 3)


def real_roots(coeffs_low_to_high):
    p = np.ascontiguousarray(coeffs_low_to_high, np.float64)
    r = np.zeros(12)
    n = host_lib().nvs_host_real_roots(p.ctypes.data, len(p) - 1, r.ctypes.data)
    return r[:n]


def rot_angle_deg(Ra, Rb):
    """Angle of Ra^T Rb via atan2(|skew part|, (trace-1)/2): well conditioned near zero (arccos of the trace alone has
    a 0.02 degree noise floor for float32 matrices)."""
    M = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    v = 0.5 * np.array([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    return float(np.degrees(np.arctan2(np.linalg.norm(v), (np.trace(M) - 1) / 2)))


def dir_angle_deg(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.degrees(np.arctan2(np.linalg.norm(np.cross(a, b)), a @ b)))


def sampson_sq(E, cur, ref):
    A = np.c_[cur, np.ones(len(cur))].astype(np.float64)
    B = np.c_[ref, np.ones(len(ref))].astype(np.float64)
    Ea, Etb = A @ np.asarray(E, np.float64).T, B @ np.asarray(E, np.float64)
    r = (B * Ea).sum(1)
    return r * r / (Ea[:, 0] ** 2 + Ea[:, 1] ** 2 + Etb[:, 0] ** 2 + Etb[:, 1] ** 2)
