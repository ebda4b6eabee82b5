"""Per-kernel parity: CUDA path (through the C ABI) vs plain torch fp32 on CPU / the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


def _ops():
    from nano_vs_slam_b200 import ops
    return ops


def _ref_conv(x, w, b, act, k):
    y = F.conv2d(x, w, b, padding=k // 2)
    if act == 1:
        y = F.leaky_relu(y, 0.01)
    elif act == 2:
        y = F.relu(y)
    elif act == 3:
        y = y.sigmoid()
    elif act == 4:
        y = y.tanh()
    elif act == 5:
        y = torch.cat([y[:, :1].sigmoid(), y[:, 1:].tanh()], 1)
    elif act == 6:
        y = F.gelu(y)
    return y


CONV_CASES = [
    # cin, cout, k, H, W, act
    (3, 16, 3, 40, 56, 1), (16, 32, 3, 24, 40, 1), (32, 32, 3, 33, 47, 1), (64, 64, 3, 30, 40, 1),
    (64, 128, 3, 15, 20, 0), (64, 1, 3, 17, 23, 3), (64, 2, 3, 17, 23, 4), (48, 3, 3, 12, 20, 5),
    (64, 28, 3, 20, 28, 0), (24, 19, 3, 20, 28, 0), (48, 48, 3, 19, 38, 1), (24, 24, 3, 16, 16, 2),
    (48, 96, 3, 9, 19, 1), (64, 64, 1, 10, 14, 0), (128, 128, 1, 10, 14, 6), (96, 48, 1, 5, 7, 0),
    (64, 64, 3, 60, 80, 1), (16, 32, 3, 64, 96, 1),
]


@pytest.mark.parametrize("cin,cout,k,H,W,act", CONV_CASES)
def test_conv_plain(cin, cout, k, H, W, act):
    ops = _ops()
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    wp, bp = ops.pack_conv(w.cuda(), bias=b.cuda())
    y = ops.conv(x.cuda(), wp, bp, cout, ksize=k, act=act)
    torch.cuda.synchronize()
    assert rel_err(y, _ref_conv(x, w, b, act, k)) < 2e-5


def test_conv_bn_fold_pool_both_shuffle_concat_slice():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 22, 38
    # conv + BN + lrelu, output full + pooled (backbone conv3b)
    x = torch.randn(B, 32, H, W, generator=g)
    conv = torch.nn.Conv2d(32, 64, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(64).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        ref = F.leaky_relu(bn(conv(x)), 0.01)
    bnd = {"weight": bn.weight.cuda(), "bias": bn.bias.cuda(), "running_mean": bn.running_mean.cuda(),
           "running_var": bn.running_var.cuda()}
    wp, bp = ops.pack_conv(conv.weight.detach().cuda(), bn=bnd)
    full, pooled = ops.conv(x.cuda(), wp, bp, 64, act=1, out_mode=ops.OUT_BOTH)
    only = ops.conv(x.cuda(), wp, bp, 64, act=1, out_mode=ops.OUT_POOL)
    assert rel_err(full, ref) < 2e-5
    assert rel_err(pooled, F.max_pool2d(ref, 2, 2)) < 2e-5
    assert torch.equal(only, pooled)
    # pixel shuffle epilogue (heads.py:98) incl. odd width
    for (hh, ww) in ((11, 19), (12, 20)):
        xs = torch.randn(B, 64, hh, ww, generator=g)
        w = torch.randn(128, 64, 3, 3, generator=g) * 0.05
        b = torch.randn(128, generator=g) * 0.1
        wp, bp = ops.pack_conv(w.cuda(), bias=b.cuda())
        y = ops.conv(xs.cuda(), wp, bp, 128, out_mode=ops.OUT_SHUFFLE)
        assert rel_err(y, F.pixel_shuffle(F.conv2d(xs, w, b, padding=1), 2)) < 2e-5
    # two-source (concat) read + channel-slice read + channel-offset write
    a = torch.randn(B, 32, H, W, generator=g)
    s = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(64, 96, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g) * 0.1
    wp, bp = ops.pack_conv(w.cuda(), bias=b.cuda())
    y = ops.conv(a.cuda(), wp, bp, 64, src1=s.cuda(), act=1)
    assert rel_err(y, F.leaky_relu(F.conv2d(torch.cat([a, s], 1), w, b, padding=1), 0.01)) < 2e-5
    w2 = torch.randn(28, 32, 3, 3, generator=g) * 0.05
    b2 = torch.randn(28, generator=g) * 0.1
    wp2, bp2 = ops.pack_conv(w2.cuda(), bias=b2.cuda())
    y2 = ops.conv(s.cuda(), wp2, bp2, 28, c0_off=32, c0=32)
    assert rel_err(y2, F.conv2d(s[:, 32:], w2, b2, padding=1)) < 2e-5


def test_conv_s2d_2x2_stride2():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    for (C, h, w) in ((64, 10, 14), (48, 5, 7)):
        x = torch.randn(2, C, h, w, generator=g)
        wt = torch.randn(2 * C, C, 2, 2, generator=g) * 0.1
        wp, bp = ops.pack_conv(wt.cuda(), s2d=True)
        y = ops.conv(x.cuda(), wp, bp, 2 * C, ksize=1, in_mode=ops.IN_S2D)
        assert rel_err(y, F.conv2d(x, wt, stride=2)) < 2e-5


def test_small_ops():
    ops = _ops()
    from oracle import kp2dtiny_ref as R
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 9, 13, generator=g)
    gg = 1 + 0.1 * torch.randn(1, 64, 1, 1, generator=g)
    bb = 0.1 * torch.randn(1, 64, 1, 1, generator=g)
    y = ops.channel_layernorm(x.cuda(), gg.reshape(-1).cuda(), bb.reshape(-1).cuda())
    assert rel_err(y, R.channel_layernorm(x, gg, bb)) < 1e-5
    assert rel_err(ops.softmax_channels(x.cuda()), x.softmax(1)) < 1e-5
    assert rel_err(ops.l2norm_channels(x.cuda()), F.normalize(x, dim=1)) < 1e-5
    w = torch.randn(64, 1, 3, 3, generator=g)
    b = torch.randn(64, generator=g)
    y = ops.dwconv3x3(x.cuda(), w.reshape(64, 9).cuda(), b.cuda())
    assert rel_err(y, F.conv2d(x, w, b, padding=1, groups=64)) < 1e-5


@pytest.mark.parametrize("C,h,w", [(64, 16, 20), (48, 10, 14), (64, 5, 7), (64, 30, 40),
                                   (256, 12, 18), (256, 30, 40)])  # head_dim 64: letters D / D_A
def test_attention(C, h, w):
    ops = _ops()
    g = torch.Generator().manual_seed(C + h)
    heads, d = 4, C // 4
    q = torch.randn(2, C, h, w, generator=g)
    kv = torch.randn(2, 2 * C, h // 2, w // 2, generator=g)
    out = ops.attention(q.cuda(), kv.cuda(), heads)

    def split(t):
        b, _, hh, ww = t.shape
        return t.reshape(b, heads, d, hh * ww).permute(0, 1, 3, 2)

    qq, kk, vv = split(q), split(kv[:, :C]), split(kv[:, C:])
    ref = torch.matmul((torch.matmul(qq, kk.transpose(-1, -2)) * d ** -0.5).softmax(-1), vv)
    ref = ref.permute(0, 1, 3, 2).reshape(2, C, h, w)
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("C,h,w", [(64, 130, 134), (48, 130, 138)])
def test_attention_large_maps_four_queries_per_thread(C, h, w):
    """Nq >= 16384 selects the 4-queries-per-thread instantiations (attention.cu, the 512x1024 S_A workload of
    BASELINE config 4); head_dim 16 and 12, ragged Nq (not a multiple of 512) and Nk (not a multiple of 256 or 8).
    The CPU reference is the literal softmax(q k^T d^-1/2) v of segformer.py:113-133, evaluated in query blocks."""
    ops = _ops()
    g = torch.Generator().manual_seed(C + h)
    heads, d = 4, C // 4
    Nq, Nk = h * w, (h // 2) * (w // 2)
    assert Nq >= 16384 and Nq % 512 != 0 and Nk % 8 != 0
    q = torch.randn(1, C, h, w, generator=g)
    kv = torch.randn(1, 2 * C, h // 2, w // 2, generator=g)
    out = ops.attention(q.cuda(), kv.cuda(), heads)

    def split(t):
        b, _, hh, ww = t.shape
        return t.reshape(b, heads, d, hh * ww).permute(0, 1, 3, 2)

    qq, kk, vv = split(q), split(kv[:, :C]), split(kv[:, C:])
    ref = torch.empty_like(qq)
    for r0 in range(0, Nq, 4096):
        sim = torch.matmul(qq[:, :, r0:r0 + 4096], kk.transpose(-1, -2)) * d ** -0.5
        ref[:, :, r0:r0 + 4096] = torch.matmul(sim.softmax(-1), vv)
    ref = ref.permute(0, 1, 3, 2).reshape(1, C, h, w)
    assert rel_err(out, ref) < 2e-5


@pytest.mark.parametrize("C,K,h,w,B", [(64, 64, 15, 20, 2), (48, 32, 10, 14, 3), (48, 64, 60, 80, 1), (64, 64, 7, 9, 5)])
def test_netvlad(C, K, h, w, B):
    ops = _ops()
    from oracle import kp2dtiny_ref as R
    g = torch.Generator().manual_seed(K + h)
    x = torch.randn(B, C, h, w, generator=g)
    sd = {"vlad_head.netvlad.centroids": torch.rand(K, C, generator=g),
          "vlad_head.netvlad.conv.weight": torch.randn(K, C, 1, 1, generator=g)}
    out = ops.netvlad(x.cuda(), sd["vlad_head.netvlad.conv.weight"].reshape(K, C).cuda(),
                      sd["vlad_head.netvlad.centroids"].cuda())
    assert rel_err(out, R.netvlad_literal(x, sd)) < 1e-4  # north_star: VLAD within 1e-4 relative


@pytest.mark.parametrize("H,W,v", [(40, 56, 2), (72, 153, 3), (240, 320, 2)])
def test_decode_and_argmax(H, W, v):
    ops = _ops()
    from oracle import kp2dtiny_ref as R
    g = torch.Generator().manual_seed(H)
    B, Hc, Wc, Hf, Wf = 2, H // 4, (W // 2) // 2, H // 2, W // 2
    score = torch.rand(B, 1, Hc, Wc, generator=g)
    shift = torch.rand(B, 2, Hc, Wc, generator=g) * 2 - 1
    feat = torch.randn(B, 32, Hf, Wf, generator=g)
    seg = torch.randn(B, 19, Hf, Wf, generator=g)
    if v == 3:
        seg = seg.softmax(1)
    a = R.Arch(v, (16, 32, 32, 64, 64, 128), n_classes=19)
    ref = R.post_processing({"score": score, "coord": shift, "feat": feat, "seg": seg, "vlad": None}, H, W, a)
    s, c, f = ops.decode(score.cuda(), shift.cuda(), feat.cuda(), H, W, 4, 2.0)
    assert torch.equal(s.cpu(), ref["score"])
    assert float((c.cpu() - ref["coord"]).abs().max()) <= 1e-3  # north_star: coordinates within 1e-3 px
    assert rel_err(f, ref["feat"]) < 1e-4
    am = ops.seg_argmax(seg.cuda())
    assert am.dtype == torch.int64 and torch.equal(am.cpu(), ref["seg"])
    ref_s = R.post_processing({"score": score, "coord": shift, "feat": feat, "seg": seg, "vlad": None}, H, W, a,
                              sample_segmentation=True)
    am_s = ops.seg_argmax(seg.cuda(), c, H, W)
    assert (am_s.cpu() == ref_s["seg"]).float().mean() >= 0.999


@pytest.mark.parametrize("n_cells,k,thr", [(4800, 1000, 0.7), (4800, 300, 0.2), (29140, 4000, 0.5), (140, 50, 0.9),
                                            (32768, 4000, 0.0)])
def test_select_keypoints(n_cells, k, thr):
    ops = _ops()
    from oracle import glue_ref
    g = torch.Generator().manual_seed(n_cells + k)
    B, D = 3, 32
    Wc = 20 if n_cells % 20 == 0 else 1
    Hc = n_cells // Wc
    score = torch.rand(B, 1, Hc, Wc, generator=g)
    score[0, 0, :3, 0] = score[0, 0, 5, 0]  # exact ties
    coord = torch.rand(B, 2, Hc, Wc, generator=g) * 100
    feat = torch.randn(B, D, Hc, Wc, generator=g)
    r = ops.select_keypoints(score.cuda(), coord.cuda(), feat.cuda(), thr, k)
    torch.cuda.synchronize()
    for b in range(B):
        post = {"score": score[b:b + 1], "coord": coord[b:b + 1], "feat": feat[b:b + 1], "seg": torch.zeros(1, 1, Hc, Wc, dtype=torch.int64)}
        pts, desc, _, cells = glue_ref.frontend_decode(post, D, thr, k)
        n = int(r["count"][b])
        assert n == len(cells)
        got = r["cell"][b, :n].cpu().numpy()
        assert np.all(np.diff(got) > 0)
        sc = score[b].reshape(-1).numpy()
        # argpartition picks ties arbitrarily: compare as score multisets, and as sets away from ties
        assert np.array_equal(np.sort(sc[got]), np.sort(sc[cells]))
        inter = len(set(got.tolist()) & set(cells.tolist()))
        assert inter >= n - 3
        np.testing.assert_array_equal(r["pts"][b, :n].cpu().numpy(), coord[b].reshape(2, -1).t().numpy()[got])
        np.testing.assert_array_equal(r["desc"][b, :n].cpu().numpy(), feat[b].reshape(D, -1).t().numpy()[got])


def test_select_semantic_filter():
    ops = _ops()
    from oracle import glue_ref
    g = torch.Generator().manual_seed(3)
    Hc, Wc, D = 30, 40, 32
    score = torch.rand(1, 1, Hc, Wc, generator=g)
    coord = torch.rand(1, 2, Hc, Wc, generator=g)
    feat = torch.randn(1, D, Hc, Wc, generator=g)
    seg = torch.randint(0, 28, (1, 1, Hc, Wc), generator=g)
    r = ops.select_keypoints(score.cuda(), coord.cuda(), feat.cuda(), 0.5, 200, seg_cells=seg.cuda(), classes_to_filter=[21, 3])
    post = {"score": score, "coord": coord, "feat": feat, "seg": seg}
    _, _, labels, cells = glue_ref.frontend_decode(post, D, 0.5, 200, classes_to_filter=[21, 3])
    n = int(r["count"][0])
    assert n == len(cells) and set(r["cell"][0, :n].cpu().tolist()) == set(cells.tolist())
    assert sorted(r["label"][0, :n].cpu().tolist()) == sorted(labels.tolist())


def test_match_against_cv2_golden_and_oracle():
    import os
    ops = _ops()
    from oracle import glue_ref
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "matcher_cv2.npz"))
    d1, d2 = torch.from_numpy(z["des1"]).cuda(), torch.from_numpy(z["des2"]).cuda()
    idx, dist = ops.match(d1, d2, mode=2)
    assert np.array_equal(idx.cpu().numpy(), z["idx"])  # vs cv2.BFMatcher.knnMatch
    np.testing.assert_allclose(dist.cpu().numpy(), z["dist"], rtol=2e-5, atol=1e-6)
    i1, i2, dd, cnt = ops.match(d1, d2, ratio=0.7, mode=0)
    r1, r2, rs = glue_ref.good_matches_one_to_one(z["idx"], z["dist"], 0.7)
    n = int(cnt)
    assert i1[:n].cpu().tolist() == r1 and i2[:n].cpu().tolist() == r2
    np.testing.assert_allclose(dd[:n].cpu().numpy(), np.array(rs, dtype=np.float32), rtol=2e-5)
    m1, m2, md, mc = ops.match(d1, d2, mode=1)
    n = int(mc)
    assert sorted(zip(m1[:n].cpu().tolist(), m2[:n].cpu().tolist())) == sorted(map(tuple, z["cross"].tolist()))


def test_match_large_random():
    ops = _ops()
    from oracle import glue_ref
    g = torch.Generator().manual_seed(1)
    n1, n2 = 4000, 3700
    b = F.normalize(torch.randn(n2, 32, generator=g), dim=1)
    src = torch.randint(0, n2, (n1,), generator=g)
    a = F.normalize(b[src] + 0.3 * torch.rand(n1, 1, generator=g) * torch.randn(n1, 32, generator=g), dim=1)
    i1, i2, dd, cnt = ops.match(a.cuda(), b.cuda(), ratio=0.7, mode=0)
    r1, r2, rs = glue_ref.bf_match(a.numpy(), b.numpy(), 0.7)
    n = int(cnt)
    # a handful of near-threshold ratio tests may flip with fp32 summation order
    got, ref = set(zip(i1[:n].cpu().tolist(), i2[:n].cpu().tolist())), set(zip(r1, r2))
    assert len(got ^ ref) <= max(2, len(ref) // 500)


@pytest.mark.parametrize("B,C,h,w,p", [(2, 48, 16, 20, 3.0), (1, 64, 8, 12, 2.6), (3, 64, 60, 80, 3.3)])
def test_gem_unshuffle4(B, C, h, w, p):
    """GeM over PixelUnshuffle(4) (aggregators/gem.py:21-33) vs torch."""
    import torch.nn.functional as F
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(B * C + h)
    x = torch.randn(B, C, h, w, generator=g)
    u = F.pixel_unshuffle(x, 4)
    ref = F.avg_pool2d(u.clamp(min=1e-6).pow(p), (u.size(-2), u.size(-1))).pow(1.0 / p).flatten(1)
    out = ops.gem(x.cuda(), p, 1e-6)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 2e-6, rel_err(out, ref)
    with pytest.raises(Exception):
        ops.gem(torch.randn(1, 8, 10, 14).cuda(), 3.0)  # PixelUnshuffle(4) needs multiples of 4


@pytest.mark.parametrize("B,C,h,w", [(2, 64, 10, 14), (1, 48, 7, 5), (3, 128, 60, 80)])
def test_convap(B, C, h, w):
    """ConvAP (aggregators/convap.py:29-37): 1x1 conv + AdaptiveAvgPool2d((4,4)) + L2 norm, ragged bins."""
    import torch.nn.functional as F
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(C + h)
    x = torch.randn(B, C, h, w, generator=g)
    wt = torch.randn(C, C, 1, 1, generator=g) * 0.2
    b = torch.randn(C, generator=g) * 0.1
    ref = F.normalize(F.adaptive_avg_pool2d(F.conv2d(x, wt, b), (4, 4)).flatten(1), p=2.0, dim=1)
    out = ops.convap(x.cuda(), wt.view(C, C).contiguous().cuda(), b.cuda(), 4, 4)
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 5e-6, rel_err(out, ref)
