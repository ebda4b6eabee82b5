"""IndexFlatL2 (fp16 tcgen05 screen with a proven error bound + fp32 re-rank + exact scan for what the screen cannot
decide) vs the fp32 oracle, vs scikit-learn golden vectors, and on data the screen alone would get wrong."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retrieval_sklearn.npz")


def _golden_cases():
    z = np.load(GOLDEN)
    return {n: (z[n + "_db"], z[n + "_q"], z[n + "_I"], z[n + "_D"], int(z[n + "_k"])) for n in ("unit", "scaled", "cluster")}


@pytest.mark.parametrize("name", ["unit", "scaled", "cluster"])
def test_oracle_is_pinned_to_sklearn_golden(name):
    """Second, independent implementation (sklearn brute force, float64) -> tests/golden/retrieval_sklearn.npz
    (oracle/gen_retrieval_golden.py): the fp32 restatement of IndexFlatL2 must return the same labels."""
    from oracle import glue_ref

    db, q, I, D, k = _golden_cases()[name]
    Dr, Ir = glue_ref.flat_l2_search(torch.from_numpy(db), torch.from_numpy(q), k)
    assert np.array_equal(Ir.numpy(), I)
    np.testing.assert_allclose(Dr.numpy(), D, rtol=2e-4, atol=2e-4 * float(D.max()))


def _exact64(db, q):
    """float64 squared distances of the fp32 data (the ground truth the fp32 formula approximates)."""
    a, b = q.double().cpu(), db.double().cpu()
    return (a * a).sum(1, keepdim=True) + (b * b).sum(1).unsqueeze(0) - 2.0 * a @ b.t()


def _check_exact(I, D, d64, k, tol):
    """Every returned row is one of the k nearest up to fp32 noise `tol`, every row clearly inside the k-th distance is
    returned, distances are the fp32 values, ascending, ties by id."""
    I, D = I.cpu(), D.cpu()
    srt = d64.sort(dim=1).values
    dk = srt[:, k - 1:k]
    got = torch.gather(d64, 1, I)
    assert bool((got <= dk + tol).all()), float((got - dk).max())
    must = d64 < dk - tol
    hit = torch.zeros_like(must)
    hit.scatter_(1, I, True)
    assert bool((hit | ~must).all()), "a row clearly inside the k-th distance is missing"
    assert bool((I >= 0).all()) and all(len(set(r.tolist())) == k for r in I)
    np.testing.assert_allclose(D.numpy(), got.numpy(), rtol=1e-4, atol=tol)
    assert bool((D[:, 1:] >= D[:, :-1]).all())


gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("n_db,nq,d,k", [
    (5000, 300, 256, 10),      # ragged: 5000 % 256 != 0, 300 % 128 != 0
    (20000, 128, 4096, 25),    # V2-S VLAD width, BASELINE k
    (9000, 77, 1536, 20),      # V2-N width, faiss call-site k
    (3000, 10, 100, 5),        # d not a multiple of 64 (padded operand)
    (70000, 256, 512, 31),     # several strips, 3-stage kernel (lists of 38)
    (40000, 130, 256, 64),     # k at the limit (lists of 79)
])
def test_flat_l2_matches_oracle(n_db, nq, d, k):
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    db, q, planted = planted_retrieval_set(n_db, nq, d, k, seed=n_db, device="cuda")
    index = IndexFlatL2(d)
    index.add(db)
    D, I = index.search(q, k)
    torch.cuda.synchronize()
    Dr, Ir = glue_ref.flat_l2_search(db.cpu(), q.cpu(), k)
    assert I.dtype == torch.int64
    assert torch.equal(I.cpu(), Ir), "top-k index lists must be bit-exact on tie-free data"
    assert torch.equal(I.cpu(), planted.cpu())
    np.testing.assert_allclose(D.cpu().numpy(), Dr.numpy(), rtol=1e-4, atol=2e-6)
    assert bool((D[:, 1:] >= D[:, :-1]).all())


@gpu
@pytest.mark.parametrize("name", ["unit", "scaled", "cluster"])
def test_flat_l2_matches_sklearn_golden(name):
    from nano_vs_slam_b200.retrieval import IndexFlatL2

    db, q, I, D, k = _golden_cases()[name]
    index = IndexFlatL2(db.shape[1])
    index.add(db)
    Dg, Ig = index.search(q, k)  # numpy in -> numpy out, like faiss
    assert np.array_equal(Ig, I)
    np.testing.assert_allclose(Dg, D, rtol=2e-4, atol=2e-4 * float(D.max()))


@gpu
@pytest.mark.parametrize("n_db,d,n_cen,jitter,contiguous,k", [
    (20000, 256, 300, 1e-3, False, 25),   # ~67 rows per cluster, all within ~2e-3 of each other: dozens of rows sit
    (20000, 256, 300, 1e-3, True, 25),    # inside the screen's slack; contiguous clusters fill one strip's list -> scan
    (12000, 4096, 150, 1e-4, True, 20),   # VLAD width, near-duplicates 1e-4 apart
    (30000, 128, 500, 3e-3, False, 40),
])
def test_flat_l2_exact_on_clustered_data(n_db, d, n_cen, jitter, contiguous, k):
    """Real descriptors are not planted: many database rows lie within the half-precision error of the k-th neighbour.
    The result must still be the fp32 result (up to fp32 noise in the distances themselves), twice in a row."""
    from nano_vs_slam_b200.retrieval import IndexFlatL2

    g = torch.Generator(device="cuda").manual_seed(n_db + d)
    cen = torch.nn.functional.normalize(torch.randn(n_cen, d, generator=g, device="cuda"), dim=1)
    assign = torch.randint(0, n_cen, (n_db,), generator=g, device="cuda")
    if contiguous:
        assign = assign.sort().values
    noise = torch.nn.functional.normalize(torch.randn(n_db, d, generator=g, device="cuda"), dim=1)
    db = cen[assign] + jitter * torch.rand(n_db, 1, generator=g, device="cuda") * noise
    nq = 200
    qn = torch.nn.functional.normalize(torch.randn(nq, d, generator=g, device="cuda"), dim=1)
    q = cen[torch.randint(0, n_cen, (nq,), generator=g, device="cuda")] + 2 * jitter * qn
    index = IndexFlatL2(d)
    index.add(db)
    D1, I1 = index.search(q, k)
    D2, I2 = index.search(q, k)
    torch.cuda.synchronize()
    assert torch.equal(I1, I2) and torch.equal(D1, D2), "results must not depend on scheduling"
    _check_exact(I1, D1, _exact64(db, q), k, tol=2e-6)


@gpu
def test_flat_l2_duplicate_rows_and_ties():
    """Exact ties (every row stored 40 times): the k nearest are the copies with the LOWEST ids (ties -> lower id)."""
    from nano_vs_slam_b200.retrieval import IndexFlatL2

    g = torch.Generator().manual_seed(5)
    base = torch.nn.functional.normalize(torch.randn(200, 128, generator=g), dim=1)
    db = base.repeat(40, 1)                      # row j == row j % 200
    q = base[:50] + 0.01 * torch.randn(50, 128, generator=g)
    k = 25
    index = IndexFlatL2(128)
    index.add(db.cuda())
    D, I = index.search(q.cuda(), k)
    d64 = _exact64(db, q)
    # expected: sort by (fp32 distance as the product computes it, id); copies have bit-identical fp32 distances, so
    # the nearest base row contributes its 25 lowest ids
    nearest = d64[:, :200].argmin(1)
    expect = nearest.unsqueeze(1) + 200 * torch.arange(k).unsqueeze(0)
    assert torch.equal(I.cpu(), expect)
    assert bool((D[:, 1:] == D[:, :1]).all())


@gpu
def test_flat_l2_nan_query_and_small_database():
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from oracle import glue_ref

    g = torch.Generator().manual_seed(6)
    db = torch.randn(40, 64, generator=g)
    q = torch.randn(6, 64, generator=g)
    q[2, 5] = float("nan")
    index = IndexFlatL2(64)
    index.add(db.cuda())
    D, I = index.search(q.cuda(), 8)
    ok = [0, 1, 3, 4, 5]
    Dr, Ir = glue_ref.flat_l2_search(db, q[ok], 8)
    assert torch.equal(I.cpu()[ok], Ir)
    assert bool((I[2] == -1).all()) and bool(torch.isinf(D[2]).all())  # no neighbour is reachable: faiss's (-1, inf)
    with pytest.raises(Exception):
        index.search(q.cuda(), 41)   # k > ntotal
    with pytest.raises(NotImplementedError):
        index.search(q.cuda(), 65)      # k > nvs_flat_max_k()


@gpu
def test_flat_l2_numpy_api_and_sharded_merge():
    from nano_vs_slam_b200.retrieval import IndexFlatL2, ShardedIndexFlatL2, merge_topk_device, shard_bounds
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    n_db, nq, d, k = 12000, 201, 256, 21     # nq * k odd: the packed all_gather buffer pads
    db, q, planted = planted_retrieval_set(n_db, nq, d, k, seed=1, device="cuda")
    index = IndexFlatL2(d)
    index.add(db.cpu().numpy())
    D, I = index.search(q.cpu().numpy(), k)          # faiss-shaped numpy call
    assert isinstance(D, np.ndarray) and I.dtype == np.int64 and np.array_equal(I, planted.cpu().numpy())
    # 8 shards searched in-process, each result written into its slot of ONE packed buffer (labels | distances), merged
    # in place exactly as ShardedIndexFlatL2 does after its single all_gather
    words = nq * k + (nq * k + 1) // 2
    gathered = torch.empty(8, words, dtype=torch.int64, device="cuda")
    Dg, Ig = ShardedIndexFlatL2._packed_views(gathered, 8, nq, k)
    for r in range(8):
        lo, hi = shard_bounds(n_db, 8, r)
        sh = IndexFlatL2(d)
        sh.add(db[lo:hi])
        sh.search_device(q, k, id_offset=lo, out=(Dg[r], Ig[r]))
    Dm, Im = merge_topk_device(Dg, Ig)
    assert torch.equal(Im, planted)
    Dr, Ir = glue_ref.flat_l2_search(db.cpu(), q.cpu(), k)
    np.testing.assert_allclose(Dm.cpu().numpy(), Dr.numpy(), rtol=1e-4, atol=2e-6)
    Dc, Ic = merge_topk_device(Dg.contiguous(), Ig.contiguous())  # plain [parts][nq][k] layout
    assert torch.equal(Ic, Im) and torch.equal(Dc, Dm)


@gpu
@pytest.mark.parametrize("shards", [2, 5])
def test_flat_l2_two_phase_sharded_search_is_exact(shards):
    """nvs_flat_search_begin / _end (the per-shard path of ShardedIndexFlatL2 on GPUs): the shards agree on a bound of the
    global k-th distance (k-th smallest of the union of their published bounds) and re-rank only what lies inside it; merged, the result is the
    single-index result -- on planted data bit for bit, on clustered data (near ties across the shards) the exact set --
    and each shard re-ranks far fewer rows than k."""
    from nano_vs_slam_b200.retrieval import IndexFlatL2, merge_bounds, merge_topk_device, shard_bounds
    from nano_vs_slam_b200.synthetic import planted_retrieval_set

    k = 25
    for kind in ("planted", "cluster"):
        if kind == "planted":
            db, q, planted = planted_retrieval_set(40000, 300, 256, k, seed=3, device="cuda")
        else:
            g = torch.Generator(device="cuda").manual_seed(9)
            cen = torch.nn.functional.normalize(torch.randn(300, 128, generator=g, device="cuda"), dim=1)
            db = cen[torch.randint(0, 300, (30000,), generator=g, device="cuda")] + 1e-3 * torch.randn(30000, 128, generator=g, device="cuda")
            q = cen[torch.randint(0, 300, (200,), generator=g, device="cuda")] + 2e-3 * torch.randn(200, 128, generator=g, device="cuda")
        n, d = db.shape
        whole = IndexFlatL2(d)
        whole.add(db)
        Dw, Iw = whole.search(q, k)
        parts, bounds = [], []
        for r in range(shards):
            lo, hi = shard_bounds(n, shards, r)
            ix = IndexFlatL2(d)
            ix.add(db[lo:hi].contiguous())
            parts.append((ix, lo))
            bounds.append(ix.search_begin(q, k))
        gb = merge_bounds(torch.stack(bounds))
        Ds, Is = [], []
        for ix, lo in parts:
            D, I = ix.search_end(q, k, gb, id_offset=lo)
            Ds.append(D)
            Is.append(I)
        Dm, Im = merge_topk_device(torch.stack(Ds), torch.stack(Is))
        torch.cuda.synchronize()
        if kind == "planted":
            assert torch.equal(Im, Iw) and torch.equal(Im.cpu(), planted.cpu())
            assert torch.equal(Dm, Dw)
        else:
            _check_exact(Im, Dm, _exact64(db, q), k, tol=2e-6)
        # the point of the exchange: on data with distinct neighbours the shards together re-rank about k rows per
        # query, not k each (clustered data: ~100 rows per query are inside the error bound wherever they live)
        kept = sum(int((I >= 0).sum()) for I in Is) / (q.shape[0] * k)
        assert kind == "cluster" or kept < 1.3, kept


def test_query_chunks_cover_the_queries():
    """Host logic of the pipelined sharded search: contiguous chunks, a short last chunk joins its predecessor."""
    from nano_vs_slam_b200.retrieval import ShardedIndexFlatL2

    for nq, chunk in ((10000, 4096), (4096, 4096), (4097, 4096), (4700, 4096), (1, 4096), (9000, 1024), (300, 128)):
        ch = ShardedIndexFlatL2.query_chunks(nq, chunk)
        assert ch[0][0] == 0 and ch[-1][1] == nq
        assert all(a < b for a, b in ch) and all(ch[i][1] == ch[i + 1][0] for i in range(len(ch) - 1))
        assert all(b - a <= chunk + 511 for a, b in ch)
        assert len(ch) == 1 or ch[-1][1] - ch[-1][0] >= min(512, chunk)
    assert ShardedIndexFlatL2.query_chunks(10000, 4096) == [(0, 4096), (4096, 8192), (8192, 10000)]
    assert ShardedIndexFlatL2.query_chunks(10000) == [(0, 10000)]  # default: one chunk


@gpu
def test_pipelined_sharded_search_matches_single_index(tmp_path, monkeypatch):
    """ShardedIndexFlatL2._search_pipelined (query chunks: GEMM on the main stream, exchange tail on a side stream, two
    workspace slots) in a one-rank NCCL group, five chunks: identical to the plain single-index search."""
    import torch.distributed as dist

    from nano_vs_slam_b200.retrieval import IndexFlatL2, ShardedIndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set

    created = not dist.is_initialized()
    if created:
        dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        k = 25
        db, q, planted = planted_retrieval_set(40000, 1300, 256, k, seed=5, device="cuda")
        whole = IndexFlatL2(256)
        whole.add(db)
        Dw, Iw = whole.search(q, k)
        monkeypatch.setenv("NVS_RETR_CHUNK", "256")
        idx = ShardedIndexFlatL2(256, db.shape[0], device="cuda")
        idx.add_local(db)
        assert len(idx.query_chunks(q.shape[0])) == 5  # 4 x 256 + 276
        for _ in range(2):  # second pass: the workspace slots and the side stream are reused
            ge = []
            D, I = idx._search_pipelined(q, k, gemm_events=ge)
            torch.cuda.synchronize()
            assert len(ge) == 5 and all(a.elapsed_time(b) > 0 for a, b in ge)
            assert torch.equal(I, Iw) and torch.equal(I.cpu(), planted.cpu())
            assert torch.equal(D, Dw)
    finally:
        if created:
            dist.destroy_process_group()


def test_work_split_fits_the_list_slots():
    """Host logic of the retrieval GEMM's work split (no GPU): the (query group, tile) steps are cut into one contiguous
    range per cluster, and the segment of a cluster inside one query group writes its lists to slot
    (cluster - first cluster of the group).  For every shape and every possible number of resident clusters that slot
    must stay below the slot count the C layout (nvs_flat_list_slots) sizes the candidate arrays for."""
    import ctypes
    import random

    from nano_vs_slam_b200 import _cabi

    lib = _cabi.lib()

    def kernel_slots(n_mgrp, n_tiles, ncl):  # restates the epilogue's segment walk (retrieval.cu)
        W, mx = n_mgrp * n_tiles, 0
        for c in range(ncl):
            w, we = W * c // ncl, W * (c + 1) // ncl
            while w < we:
                grp = w // n_tiles
                tb = w - grp * n_tiles
                te = min(n_tiles, tb + (we - w))
                w += te - tb
                g0 = grp * n_tiles
                c0 = g0 * ncl // W
                while W * (c0 + 1) // ncl <= g0:
                    c0 += 1
                while W * c0 // ncl > g0:
                    c0 -= 1
                assert c >= c0
                mx = max(mx, c - c0)
        return mx + 1

    rng = random.Random(3)
    shapes = [(1_000_000, 10_000), (125_000, 10_000), (300, 1), (256, 128), (257, 129), (70_000, 900), (5_000, 3_000)]
    shapes += [(rng.randint(1, 400_000), rng.randint(1, 6_000)) for _ in range(40)]
    for n_db, nq in shapes:
        cs = ctypes.c_int32(0)
        slots = lib.nvs_flat_list_slots(n_db, nq, 4096, 25, ctypes.byref(cs))
        assert slots >= 1 and cs.value in (2, 4, 8), (n_db, nq, slots, cs.value)
        n_tiles, n_mblk = -(-n_db // 256), -(-nq // 128)
        n_mgrp = -(-n_mblk // cs.value)
        most = min(148 // cs.value, n_mgrp * n_tiles)
        for ncl in sorted({1, 2, most // 2 or 1, max(1, most - 1), most}):
            assert kernel_slots(n_mgrp, n_tiles, ncl) <= slots, (n_db, nq, cs.value, ncl, slots)
    assert lib.nvs_flat_list_slots(0, 1, 4096, 25, None) == 0
