"""torch.ops.nanovs.* (nano_vs_slam_b200/torch_ops.py): registration, fake (meta) shapes, CUDA-only dispatch, and -- on the
GPU -- direct calls compared with the same references the ops tests use (the PyTorch custom-op boundary
serving kp2dtiny.py:552-647, :906-1015, frontend.py:94-126, feature_matcher.py:89-98, global_descriptor.py:55-60)."""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

from util import rel_err

EXPECTED = {"kp2dtiny_forward", "decode", "seg_argmax", "select_keypoints", "match", "match_batch", "pose_batch",
            "pose_batch_adaptive", "flat_l2_search", "flat_l2_begin", "flat_l2_end", "topk_merge", "conv_tc", "conv_rs", "split16", "unsplit16", "conv",
            "attention", "netvlad", "channel_layernorm", "dwconv3x3", "softmax_channels", "preprocess_u8"}


def _model(letter="S", ncls=28, v3=False):
    from nano_vs_slam_b200 import tiny_factory

    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(letter, ncls, v3=v3)
    m.eval()
    m.training = False
    return m


def test_every_op_is_registered_for_cuda_only():
    from nano_vs_slam_b200 import torch_ops

    assert set(torch_ops.OP_NAMES) == EXPECTED
    for name in EXPECTED:
        op = getattr(torch.ops.nanovs, name)
        assert torch._C._dispatch_has_kernel_for_dispatch_key(f"nanovs::{name}", "CUDA"), name
        assert not torch._C._dispatch_has_kernel_for_dispatch_key(f"nanovs::{name}", "CPU"), name  # no CPU fallback
        assert op is not None


def test_cpu_tensors_raise_from_the_dispatcher():
    import nano_vs_slam_b200.torch_ops  # noqa: F401

    with pytest.raises(NotImplementedError):
        torch.ops.nanovs.decode(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4), None, 16, 16, 4, 2.0)
    with pytest.raises(NotImplementedError):
        torch.ops.nanovs.attention(torch.zeros(1, 64, 4, 4), torch.zeros(1, 128, 2, 2), 4)
    with pytest.raises(NotImplementedError):
        torch.ops.nanovs.match(torch.zeros(8, 32), torch.zeros(8, 32), 0.7, 0)


@pytest.mark.parametrize("letter,v3,ncls,vdim", [("S", False, 28, 4096), ("N", True, 19, 64 * 48), ("S_A", False, 19, 4096)])
def test_fake_shapes_follow_the_reference_forward_dict(letter, v3, ncls, vdim):
    """Shape propagation without a GPU: the meta implementations must give the shapes of kp2dtiny.py:552-591."""
    from torch._subclasses.fake_tensor import FakeTensorMode

    from nano_vs_slam_b200 import torch_ops

    m = _model(letter, ncls, v3)
    h = torch_ops.register_model(m)
    with FakeTensorMode():
        x = torch.empty(3, 3, 240, 320, device="cuda")
        score, coord, feat, vlad, seg = torch.ops.nanovs.kp2dtiny_forward(x, h, False)
        assert score.shape == (3, 1, 60, 80) and coord.shape == (3, 2, 60, 80)
        assert feat.shape == (3, 32, 120, 160) and seg.shape == (3, ncls, 120, 160) and vlad.shape == (3, vdim)
        s, c, f = torch.ops.nanovs.decode(score, coord, feat, 240, 320, 4, 2.0)
        assert f.shape == (3, 32, 60, 80) and s.shape == score.shape and c.shape == coord.shape
        lab = torch.ops.nanovs.seg_argmax(seg, None, 240, 320)
        assert lab.shape == (3, 1, 120, 160) and lab.dtype == torch.int64
        pts, desc, sc, cell, label, count = torch.ops.nanovs.select_keypoints(s, c, f, None, [], 0.7, 1000)
        assert pts.shape == (3, 1000, 2) and desc.shape == (3, 1000, 32) and count.shape == (3,)
        assert cell.dtype == torch.int32 and count.dtype == torch.int32 and label.numel() == 0
        u8 = torch.empty(3, 376, 1241, 3, device="cuda", dtype=torch.uint8)
        assert torch.ops.nanovs.preprocess_u8(u8, [240, 320]).shape == (3, 3, 240, 320)


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_forward_and_post_processing_run_through_the_ops():
    """model(x) is torch.ops.nanovs.kp2dtiny_forward; calling the op directly gives the same tensors."""
    from nano_vs_slam_b200 import torch_ops
    from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
    from oracle import kp2dtiny_ref as R

    m = _model()
    sd = spread_init(m.state_dict(), 1234)
    m.load_state_dict(sd)
    m.eval(); m.training = False
    m = m.cuda()
    x = synthetic_frames(2, 64, 96, 0)
    out = m(x.cuda())
    direct = torch.ops.nanovs.kp2dtiny_forward(x.cuda(), torch_ops.register_model(m), False)
    ref = R.forward(x, sd, R.arch_for("S", False, 28))
    for k, t in zip(("score", "coord", "feat", "vlad", "seg"), direct):
        assert rel_err(t, ref[k]) < 1e-4, k
        assert rel_err(out[k], t) < 2e-5, k
    s, c, f = torch.ops.nanovs.decode(out["score"], out["coord"], out["feat"], 64, 96, 4, 2.0)
    rpost = R.post_processing(dict(ref), 64, 96, R.arch_for("S", False, 28))
    assert float((c.cpu() - rpost["coord"]).abs().max()) < 1e-3
    assert rel_err(f, rpost["feat"]) < 1e-4
    lab = torch.ops.nanovs.seg_argmax(out["seg"], None, 64, 96)
    assert float((lab.cpu() == rpost["seg"]).float().mean()) >= 0.999
    with pytest.raises(NotImplementedError):
        torch.ops.nanovs.kp2dtiny_forward(x, torch_ops.register_model(m), False)  # CPU tensor: dispatcher refuses


@pytest.mark.gpu
def test_unit_range_input_matches_the_normalised_input():
    """forward(x01, unit_input=True) applies x.sub(0.5).mul(2.0) (frontend.py:79) in the stem kernel's load: bit-equal
    arithmetic, so the outputs must agree with forward(x01 * 2 - 1 computed the reference's way)."""
    from nano_vs_slam_b200.synthetic import spread_init

    m = _model()
    m.load_state_dict(spread_init(m.state_dict(), 5))
    m.eval(); m.training = False
    m = m.cuda()
    g = torch.Generator().manual_seed(3)
    x01 = torch.rand(2, 3, 64, 96, generator=g).cuda()
    a = m(x01.sub(0.5).mul(2.0))
    b = m(x01, unit_input=True)
    for k in ("score", "coord", "feat", "vlad", "seg"):
        assert rel_err(b[k], a[k]) < 2e-5, k  # run-to-run rounding of the multi-issuer schedule only


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,pool", [(64, 64, False), (32, 32, True), (96, 64, False), (64, 128, False)])
def test_conv_tc_op_matches_fp32_conv(cin, cout, pool):
    """torch.ops.nanovs.conv_tc vs F.conv2d in fp32 (modules/base.py:39-46 with BN folded)."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, 24, 40, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    hi, lo, bp = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
    dst, pooled = torch.ops.nanovs.conv_tc(x_nhwc, None, hi, lo, bp, cout, ops.ACT_LRELU, 1, 1, pool)
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    assert rel_err(dst, ref) < 2e-5
    if pool:
        assert rel_err(pooled.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5
    else:
        assert pooled.numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,pool", [(64, 64, False), (32, 32, True), (96, 64, False), (16, 32, True)])
def test_conv_rs_op_matches_fp32_conv(cin, cout, pool):
    """torch.ops.nanovs.conv_rs (3xFP16, split-format activations) vs F.conv2d in fp32."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, 24, 40, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    hi, lo, bp, scale, _ = ops.pack_conv_rs(w.cuda(), bias=b.cuda()).slices[0]
    xs = torch.ops.nanovs.split16(x.permute(0, 2, 3, 1).contiguous().cuda())
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    dst, pooled = torch.ops.nanovs.conv_rs(xs, None, hi, lo, bp, scale, cout, ops.ACT_LRELU, 1, 1, pool)
    assert rel_err(dst, ref) < 2e-5
    dst2, _ = torch.ops.nanovs.conv_rs(xs, None, hi, lo, bp, scale, cout, ops.ACT_LRELU, 1, 0, False)
    assert rel_err(torch.ops.nanovs.unsplit16(dst2).permute(0, 3, 1, 2), ref) < 2e-5
    if pool:
        assert rel_err(torch.ops.nanovs.unsplit16(pooled).permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5
    else:
        assert pooled.numel() == 0
    with pytest.raises(NotImplementedError):
        torch.ops.nanovs.split16(torch.zeros(1, 4, 4, 8))


@pytest.mark.gpu
def test_small_ops_through_torch_ops():
    from oracle import kp2dtiny_ref as R

    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 9, 13, generator=g)
    assert rel_err(torch.ops.nanovs.softmax_channels(x.cuda()), x.softmax(1)) < 1e-5
    gg, bb = 1 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    y = torch.ops.nanovs.channel_layernorm(x.cuda(), gg.cuda(), bb.cuda(), 1e-5)
    assert rel_err(y, R.channel_layernorm(x, gg.view(1, -1, 1, 1), bb.view(1, -1, 1, 1))) < 1e-5
    sd = {"vlad_head.netvlad.centroids": torch.rand(64, 64, generator=g),
          "vlad_head.netvlad.conv.weight": torch.randn(64, 64, 1, 1, generator=g)}
    v = torch.ops.nanovs.netvlad(x.cuda(), sd["vlad_head.netvlad.conv.weight"].reshape(64, 64).cuda(),
                                 sd["vlad_head.netvlad.centroids"].cuda())
    assert rel_err(v, R.netvlad_literal(x, sd)) < 1e-4


@pytest.mark.gpu
def test_match_and_retrieval_ops():
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from nano_vs_slam_b200 import torch_ops
    from oracle import glue_ref

    g = torch.Generator().manual_seed(4)
    d1 = F.normalize(torch.randn(300, 32, generator=g), dim=1)
    d2 = F.normalize(d1[torch.randperm(300, generator=g)[:250]] + 0.05 * torch.randn(250, 32, generator=g), dim=1)
    i1, i2, dd, cnt = torch.ops.nanovs.match(d1.cuda(), d2.cuda(), 0.7, 0)
    r1, r2, rs = glue_ref.bf_match(d1.numpy(), d2.numpy(), 0.7)
    n = int(cnt)
    assert i1[:n].cpu().tolist() == list(r1) and i2[:n].cpu().tolist() == list(r2)
    db, q, planted = planted_retrieval_set(6000, 100, 256, 10, seed=9, device="cuda")
    index = IndexFlatL2(256)
    index.add(db)
    D, I = torch.ops.nanovs.flat_l2_search(q, torch_ops.register_index(index), 10, 0)
    assert torch.equal(I, planted)
