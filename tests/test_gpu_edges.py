"""Edge cases of the path: empty / tiny / ragged inputs, k larger than what exists, degenerate matches."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_select_no_keypoint_and_all_keypoints():
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(0)
    score = torch.rand(2, 1, 6, 10, generator=g) * 0.5
    coord = torch.rand(2, 2, 6, 10, generator=g)
    feat = torch.randn(2, 32, 6, 10, generator=g)
    r = ops.select_keypoints(score.cuda(), coord.cuda(), feat.cuda(), 0.9, 100)   # nothing passes
    assert r["count"].cpu().tolist() == [0, 0]
    r = ops.select_keypoints(score.cuda(), coord.cuda(), feat.cuda(), -1.0, 1000)  # everything passes, k > cells
    assert r["count"].cpu().tolist() == [60, 60]
    assert r["cell"][0, :60].cpu().tolist() == list(range(60))
    score[1] = 0.25  # a frame where every score ties
    r = ops.select_keypoints(score.cuda(), coord.cuda(), feat.cuda(), 0.1, 7)
    assert r["count"].cpu().tolist() == [7, 7]


def test_match_tiny_and_degenerate_sets():
    from nano_vs_slam_b200 import ops
    from oracle import glue_ref

    g = torch.Generator().manual_seed(2)
    b = F.normalize(torch.randn(2, 32, generator=g), dim=1)   # the smallest train set a 2-NN ratio test can use
    a = F.normalize(b[:1] + 0.01 * torch.randn(1, 32, generator=g), dim=1)
    i1, i2, dd, cnt = ops.match(a.cuda(), b.cuda(), ratio=0.7, mode=0)
    r1, r2, _ = glue_ref.bf_match(a.numpy(), b.numpy(), 0.7)
    assert int(cnt) == len(r1) == 1 and i2[:1].cpu().tolist() == r2
    # every query equidistant from both train descriptors: the ratio test rejects all of them
    b2 = torch.eye(32)[:2]
    a2 = F.normalize(b2.sum(0, keepdim=True).repeat(5, 1), dim=1)
    _, _, _, cnt = ops.match(a2.cuda(), b2.cuda(), ratio=0.7, mode=0)
    assert int(cnt) == 0 == len(glue_ref.bf_match(a2.numpy(), b2.numpy(), 0.7)[0])
    # many queries claiming one train descriptor: one-to-one keeps a single pair (feature_matcher.py:179-209)
    b3 = F.normalize(torch.randn(50, 32, generator=g), dim=1)
    a3 = F.normalize(b3[7:8] + 0.02 * torch.randn(20, 32, generator=g), dim=1)
    i1, i2, dd, cnt = ops.match(a3.cuda(), b3.cuda(), ratio=0.7, mode=0)
    r1, r2, _ = glue_ref.bf_match(a3.numpy(), b3.numpy(), 0.7)
    n = int(cnt)
    assert n == len(r1) == 1 and i1[:n].cpu().tolist() == r1 and i2[:n].cpu().tolist() == r2 == [7]
    m1, m2, _, mc = ops.match(a3.cuda(), b3.cuda(), mode=1)  # mutual NN: exactly one pair as well
    e1, e2, _ = glue_ref.mutual_nn(a3.numpy(), b3.numpy())
    assert int(mc) == len(e1) and sorted(m2[:int(mc)].cpu().tolist()) == sorted(e2.tolist())


def test_retrieval_single_query_small_db_and_incremental_add():
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    db, q, planted = planted_retrieval_set(64, 1, 128, 5, seed=3, device="cuda")   # one query, db smaller than a tile
    idx = IndexFlatL2(128)
    idx.add(db[:40])
    idx.add(db[40:])            # faiss-style incremental add
    assert idx.ntotal == 64
    D, I = idx.search(q, 5)
    Dr, Ir = glue_ref.flat_l2_search(db.cpu(), q.cpu(), 5)
    assert torch.equal(I.cpu(), Ir) and torch.equal(I.cpu(), planted.cpu())
    np.testing.assert_allclose(D.cpu().numpy(), Dr.numpy(), rtol=1e-4, atol=2e-6)
    with pytest.raises(NotImplementedError):
        idx.search(q, 65)       # k above nvs_flat_max_k() (the per-row lists live in shared memory)
    for kk in (31, 64):         # 64 == ntotal: every row is returned, in order
        Dk, Ik = idx.search(q, kk)
        assert torch.equal(Ik.cpu(), glue_ref.flat_l2_search(db.cpu(), q.cpu(), kk)[1])


def test_model_smallest_and_ragged_frames():
    """8x8 is the smallest legal frame (skip-level map 4x4; the attention heads need 16x16: their 2x2 / stride-2
    key-value conv runs on the H/8 map, and fails on a 1x1 map in the reference as well); 40x72 has ragged tiles."""
    import contextlib
    import io

    from nano_vs_slam_b200 import tiny_factory
    from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
    from oracle import kp2dtiny_ref as R
    from util import rel_err, tol

    for letter, v3 in (("S", False), ("N_A", True)):
        with contextlib.redirect_stdout(io.StringIO()):
            m = tiny_factory(letter, 19, v3=v3)
        sd = spread_init(m.state_dict(), 77)
        m.load_state_dict(sd)
        m.eval()
        m.training = False
        m = m.cuda()
        a = R.arch_for(letter, v3, 19)
        for (B, H, W) in ((1, 16, 16) if letter.endswith("_A") else (1, 8, 8), (3, 40, 72)):
            x = synthetic_frames(B, H, W, 5)
            out = m(x.cuda())
            ref = R.forward(x, sd, a)
            for k in ("score", "coord", "feat", "vlad", "seg"):
                assert out[k].shape == ref[k].shape
                assert rel_err(out[k], ref[k]) < tol(k, v3), (letter, H, W, k, rel_err(out[k], ref[k]))
        with pytest.raises(RuntimeError):
            m(synthetic_frames(1, 36, 36, 0).cuda())   # floor(H/2) % 4 != 0: the reference fails in torch.cat


def test_match_batch_equals_per_pair_calls():
    """nvs_match_batch (all pairs, device-side counts) against nvs_match pair by pair, incl. empty / tiny frames."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(4)
    F_, kmax, D = 6, 300, 32
    base = F.normalize(torch.randn(kmax, D, generator=g), dim=1)
    desc = torch.stack([F.normalize(base + 0.15 * f * torch.randn(kmax, D, generator=g), dim=1) for f in range(F_)]).cuda()
    counts = torch.tensor([300, 257, 1, 0, 129, 300], dtype=torch.int32).cuda()
    pa = torch.tensor([1, 0, 4, 5, 2, 3, 0], dtype=torch.int32).cuda()
    pb = torch.tensor([0, 1, 5, 4, 1, 0, 2], dtype=torch.int32).cuda()
    for mode in (0, 1):
        i1, i2, dd, cnt = ops.match_batch(desc, counts, pa, pb, ratio=0.8, mode=mode)
        for p in range(pa.numel()):
            a, b = int(pa[p]), int(pb[p])
            n1, n2 = int(counts[a]), int(counts[b])
            n = int(cnt[p])
            if n1 < 1 or n2 < (2 if mode == 0 else 1):
                assert n == 0, (mode, p, n)
                continue
            r1, r2, rd, rc = ops.match(desc[a, :n1].contiguous(), desc[b, :n2].contiguous(), ratio=0.8, mode=mode)
            assert n == int(rc), (mode, p)
            assert torch.equal(i1[p, :n], r1[:n]) and torch.equal(i2[p, :n], r2[:n]) and torch.equal(dd[p, :n], rd[:n])


@pytest.mark.parametrize("letter,v3,shape", [("S", False, (2, 64, 96)), ("N", True, (2, 64, 96)), ("N_A", False, (2, 64, 96)),
                                             ("F", False, (2, 64, 96)), ("S", False, (1, 376, 1241)),
                                             ("N", True, (1, 184, 328))])
def test_no_dependence_on_uninitialised_buffers(letter, v3, shape):
    """Plan buffers come from torch.empty: fill the allocator's cache with NaNs first, the outputs must not change
    (zero-padded channel rows of the N letters, pooled / shuffled intermediates ... are all fully written)."""
    import contextlib
    import io

    from nano_vs_slam_b200 import tiny_factory
    from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames

    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(letter, 19, v3=v3)
    m.load_state_dict(spread_init(m.state_dict(), 5))
    m.eval()
    m.training = False
    m = m.cuda()
    x = synthetic_frames(*shape, 1).cuda()
    ref = {k: v.clone() for k, v in m(x).items()}
    m._plans.clear()
    torch.cuda.synchronize()
    # NaNs into the caching allocator's large AND small pools (plan buffers of this size are mostly < 1 MiB)
    junk = [torch.full((64 << 20,), float("nan"), device="cuda") for _ in range(12)]
    junk += [torch.full((n,), float("nan"), device="cuda") for n in (1 << 10, 8 << 10, 64 << 10, 200 << 10) for _ in range(400)]
    del junk
    out = m(x)
    for k in ("score", "coord", "feat", "vlad", "seg"):
        assert torch.isfinite(out[k]).all(), k
        assert float((out[k] - ref[k]).abs().max()) <= 2e-5 * float(ref[k].abs().max()), k
