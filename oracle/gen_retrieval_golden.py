"""Golden vectors for IndexFlatL2 parity from a SECOND, independent implementation (TEST INFRASTRUCTURE).

faiss-cpu 1.9.0 (the reference's dependency, global_descriptor.py:55-60) is not installable offline, so the retrieval
oracle (oracle/glue_ref.flat_l2_search, a restatement of IndexFlatL2's published semantics) is pinned against
scikit-learn's brute-force neighbour search instead (SURVEY §8(c) probed that the two agree):

    python oracle/gen_retrieval_golden.py      # writes tests/golden/retrieval_sklearn.npz

Three small cases are stored WITH their data (so nothing depends on a random generator's stream): unit-norm descriptors
with planted neighbours, un-normalised Gaussian rows of very different lengths, and clustered rows (80 centres plus
jitter) whose neighbour gaps are of order 1e-3.  Labels are int64 row indices, distances squared L2 in float64.
"""
import os

import numpy as np
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "retrieval_sklearn.npz")


def case(name, db, q, k):
    nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(db.astype(np.float64))
    dist, idx = nn.kneighbors(q.astype(np.float64))
    # only keep queries whose k+1 nearest are separated by more than fp32 noise, so every implementation must agree
    nn2 = NearestNeighbors(n_neighbors=k + 1, algorithm="brute", metric="euclidean").fit(db.astype(np.float64))
    d2 = nn2.kneighbors(q.astype(np.float64))[0] ** 2
    ok = (np.diff(d2, axis=1).min(axis=1) > 2e-5 * np.maximum(1.0, d2[:, -1]))
    return {f"{name}_db": db.astype(np.float32), f"{name}_q": q[ok].astype(np.float32),
            f"{name}_I": idx[ok].astype(np.int64), f"{name}_D": (dist[ok] ** 2).astype(np.float64),
            f"{name}_k": np.int64(k)}


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # 1) unit descriptors, planted neighbours
    d, n, nq, k = 96, 1500, 48, 10
    db = rng.standard_normal((n, d)); db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = rng.standard_normal((nq, d)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    for i in range(nq):
        for m in range(k):
            a = 0.95 - 0.04 * m
            r = rng.standard_normal(d); r -= r.dot(q[i]) * q[i]; r /= np.linalg.norm(r)
            db[(i * k + m) * 5 % n] = a * q[i] + np.sqrt(1 - a * a) * r
    out.update(case("unit", db.astype(np.float32), q.astype(np.float32), k))
    # 2) rows of very different lengths (no normalisation), d not a multiple of 64
    d, n, nq, k = 100, 1200, 40, 7
    db = rng.standard_normal((n, d)) * rng.uniform(0.05, 20.0, (n, 1))
    q = rng.standard_normal((nq, d)) * rng.uniform(0.05, 20.0, (nq, 1))
    out.update(case("scaled", db.astype(np.float32), q.astype(np.float32), k))
    # 3) clustered rows: 80 centres + jitter, queries next to centres
    d, n, nq, k = 64, 2000, 60, 12
    cen = rng.standard_normal((80, d)); cen /= np.linalg.norm(cen, axis=1, keepdims=True)
    db = cen[rng.integers(0, 80, n)] + 0.3 * rng.standard_normal((n, d)) / np.sqrt(d)
    q = cen[rng.integers(0, 80, nq)] + 0.3 * rng.standard_normal((nq, d)) / np.sqrt(d)
    out.update(case("cluster", db.astype(np.float32), q.astype(np.float32), k))
    np.savez_compressed(OUT, **out)
    print({k_: (v.shape if hasattr(v, "shape") else v) for k_, v in out.items()})


if __name__ == "__main__":
    main()
