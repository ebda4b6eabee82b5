"""The "3xFP16" row-stationary tensor-core conv (csrc/conv_rs.cu, NvsConvTcArgs.flags bit 4) vs torch fp32 / fp64 on CPU:
every epilogue mode the launch plans use, ragged sizes, padded channels, two sources, channel slices, wide layers.
Channels-last operands are in the split fp16 hi / lo format (ops.split16 / ops.unsplit16)."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _s(t):
    """NCHW fp32 (CPU) -> the split channels-last device tensor the 3xFP16 convs read."""
    from nano_vs_slam_b200 import ops
    return ops.split16(_nhwc(t).cuda())


def _u(t):
    """Split channels-last output -> NCHW fp32."""
    from nano_vs_slam_b200 import ops
    return ops.unsplit16(t).permute(0, 3, 1, 2)


def _ref(x, w, b, act):
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).float()
    return F.leaky_relu(ref, 0.01) if act == 1 else (F.relu(ref) if act == 2 else ref)


@pytest.mark.parametrize("cin,cout,H,W,act", [
    (64, 64, 60, 80, 1), (32, 32, 24, 40, 1), (64, 128, 15, 20, 0), (96, 64, 33, 50, 1), (32, 64, 17, 31, 2),
    (64, 28, 20, 28, 0), (64, 32, 9, 19, 0), (128, 64, 8, 16, 1),
    (16, 32, 40, 56, 1),   # the stem's second conv: a 16-channel source read as one half-empty chunk
    (16, 24, 21, 37, 0),
    (64, 64, 7, 5, 1),     # tile larger than the image
    (96, 96, 11, 61, 1),   # 96 outputs = 64 + 32 slices, weights streamed (9 tiles do not fit)
    (32, 256, 6, 9, 1),    # four slices
])
def test_conv_rs_plain_nhwc_and_nchw(cin, cout, H, W, act):
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(cin + cout + H)
    B = 3
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = _ref(x, w, b, act)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="f16")
    assert isinstance(packed, ops.RsPacked)
    xs = _s(x)
    if cout % 8 == 0:
        cp = (cout + 31) // 32 * 32  # channels-last buffers are padded to 32 channels; the padding is written as zeros
        out = torch.full((B, H, W, cp), 3.0, device="cuda")
        ops.tc_conv(xs, packed, cp, act=act, dst=out, dst_layout=0).run()
        torch.cuda.synchronize()
        assert rel_err(_u(out)[:, :cout], ref) < 2e-5, rel_err(_u(out)[:, :cout], ref)
        assert cp == cout or float(_u(out)[:, cout:].abs().max()) == 0.0
    out2 = torch.zeros(B, cout, H, W, device="cuda")
    op = ops.tc_conv(xs, packed, cout, act=act, dst=None, dst_layout=1, dst_c_total=cout)
    op.run(dst_override=out2)
    torch.cuda.synchronize()
    assert rel_err(out2, ref) < 2e-5, rel_err(out2, ref)
    # three MMA issuers: the accumulation order may differ run to run in the last ulp
    out3 = torch.zeros_like(out2)
    op.run(dst_override=out3)
    assert rel_err(out3, out2) < 2e-6
    assert ops.conv_rs_range_flag() == 0


def test_conv_rs_wide_dynamic_range():
    """Activations and weights spanning many binades (the remainder parts are far below fp16's normal range for the
    small ones): the error stays at fp32 level relative to the map's maximum."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, cin, cout, H, W = 2, 64, 64, 19, 33
    x = torch.randn(B, cin, H, W, generator=g) * torch.exp(torch.randn(B, cin, H, W, generator=g) * 2)
    assert float(x.abs().max()) < 60000
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05 * torch.exp(torch.randn(cout, cin, 3, 3, generator=g) * 2)
    b = torch.randn(cout, generator=g)
    ref = _ref(x, w, b, 0)
    out = torch.zeros(B, cout, H, W, device="cuda")
    ops.tc_conv(_s(x), ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="f16"), cout, dst=out, dst_layout=1).run()
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)
    # tiny weights: the per-layer power-of-two scale keeps both parts normal
    w2 = w * 1e-6
    ref2 = _ref(x, w2, b * 1e-6, 0)
    ops.tc_conv(_s(x), ops.pack_conv_tc(w2.cuda(), bias=(b * 1e-6).cuda(), math="f16"), cout, dst=out,
                dst_layout=1).run()
    assert rel_err(out, ref2) < 2e-5, rel_err(out, ref2)


def test_conv_rs_range_flag():
    """An output beyond the fp16 range written in an activation layout raises the sticky device flag."""
    from nano_vs_slam_b200 import ops

    ops.conv_rs_range_flag(reset=True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 32, 8, 40, generator=g) * 2000
    w = torch.randn(32, 32, 3, 3, generator=g)
    out = torch.zeros(1, 8, 40, 32, device="cuda")
    ops.tc_conv(_s(x), ops.pack_conv_tc(w.cuda(), bias=torch.zeros(32).cuda(), math="f16"), 32, dst=out).run()
    assert ops.conv_rs_range_flag(reset=True) == 1  # (the fp16 a_hi of such a value is infinite)
    assert ops.conv_rs_range_flag() == 0


def test_conv_rs_pool_shuffle_concat_slice():
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 22, 38
    x = torch.randn(B, 32, H, W, generator=g)
    w = torch.randn(64, 32, 3, 3, generator=g) * 0.08
    b = torch.randn(64, generator=g) * 0.1
    ref = _ref(x, w, b, 1)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="f16")
    full = torch.zeros(B, H, W, 64, device="cuda")
    pooled = torch.zeros(B, H // 2, W // 2, 64, device="cuda")
    ops.tc_conv(_s(x), packed, 64, act=1, dst=full, dst_pool=pooled).run()
    assert rel_err(_u(full), ref) < 2e-5
    assert rel_err(_u(pooled), F.max_pool2d(ref, 2, 2)) < 2e-5
    only = torch.zeros_like(pooled)
    ops.tc_conv(_s(x), packed, 64, act=1, dst=None, dst_mode=0, dst_pool=only).run()
    assert rel_err(_u(only), _u(pooled)) < 2e-6
    # pooled output of a 32-channel layer, odd sizes (the last row / column has no partner)
    for (hh, ww) in ((21, 37), (30, 64), (8, 60)):
        xs = torch.randn(B, 16, hh, ww, generator=g)
        w1 = torch.randn(32, 16, 3, 3, generator=g) * 0.1
        b1 = torch.randn(32, generator=g) * 0.1
        p1 = torch.zeros(B, hh // 2, ww // 2, 32, device="cuda")
        ops.tc_conv(_s(xs), ops.pack_conv_tc(w1.cuda(), bias=b1.cuda(), math="f16"), 32, act=1, dst=None,
                    dst_mode=0, dst_pool=p1).run()
        assert rel_err(_u(p1), F.max_pool2d(_ref(xs, w1, b1, 1), 2, 2)) < 2e-5, (hh, ww)
    # pixel shuffle (odd and even sizes), 128 channels = two slices
    for (hh, ww) in ((11, 19), (16, 32)):
        xs = torch.randn(B, 64, hh, ww, generator=g)
        w2 = torch.randn(128, 64, 3, 3, generator=g) * 0.05
        b2 = torch.randn(128, generator=g) * 0.1
        out = torch.zeros(B, 2 * hh, 2 * ww, 32, device="cuda")
        ops.tc_conv(_s(xs), ops.pack_conv_tc(w2.cuda(), bias=b2.cuda(), math="f16"), 128, dst=out, dst_mode=2).run()
        assert rel_err(_u(out), F.pixel_shuffle(_ref(xs, w2, b2, 0), 2)) < 2e-5
    # two sources (concat) + BN fold
    a = torch.randn(B, 32, H, W, generator=g)
    s = torch.randn(B, 64, H, W, generator=g)
    conv = torch.nn.Conv2d(96, 64, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(64).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        ref = F.leaky_relu(bn(conv(torch.cat([a, s], 1))), 0.01)
    bnd = {k: getattr(bn, k).cuda() for k in ("weight", "bias", "running_mean", "running_var")}
    out = torch.zeros(B, H, W, 64, device="cuda")
    ops.tc_conv(_s(a), ops.pack_conv_tc(conv.weight.detach().cuda(), bn=bnd, math="f16"), 64, act=1,
                src1=_s(s), dst=out).run()
    assert rel_err(_u(out), ref) < 2e-5
    # channel window of a wider source
    w3 = torch.randn(32, 32, 3, 3, generator=g) * 0.08
    b3 = torch.randn(32, generator=g) * 0.1
    out = torch.zeros(B, 32, H, W, device="cuda")
    ops.tc_conv(_s(s), ops.pack_conv_tc(w3.cuda(), bias=b3.cuda(), math="f16"), 32, c0_off=32, c0=32, dst=out,
                dst_layout=1).run()
    assert rel_err(out, _ref(s[:, 32:], w3, b3, 0)) < 2e-5


def test_conv_rs_padded_channels_skip_ksteps():
    """N letters: 24 / 48 / 72 real channels in 32-channel rows (zero padding, zero weights): all-padding k-steps of 16
    channels are skipped, results unchanged."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(9)
    B, H, W = 2, 13, 35
    for real, padded, cout in ((24, 32, 24), (48, 64, 48), (72, 96, 72), (8, 32, 40)):
        x = torch.randn(B, real, H, W, generator=g)
        w = torch.randn(cout, real, 3, 3, generator=g) * 0.1
        b = torch.randn(cout, generator=g) * 0.1
        xp = torch.zeros(B, H, W, padded)
        xp[..., :real] = _nhwc(x)
        cp = (cout + 31) // 32 * 32
        out = torch.zeros(B, H, W, cp, device="cuda")
        packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), cin_segments=[(real, padded)], math="f16")
        ops.tc_conv(ops.split16(xp.cuda()), packed, cp, act=1, dst=out).run()
        got = ops.unsplit16(out)
        assert rel_err(got[..., :cout].permute(0, 3, 1, 2), _ref(x, w, b, 1)) < 2e-5, (real, cout)
        assert float(got[..., cout:].abs().max()) == 0.0 if cp > cout else True


def test_conv_rs_keypoint_heads_and_sigmoid():
    """The fused keypoint-head conv (score | location trunks -> 3 channels, sigmoid / tanh split) and a one-channel
    sigmoid (depth) output."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(77)
    B, C, H, W = 2, 64, 15, 47
    sh, lh = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ws, bs = torch.randn(1, C, 3, 3, generator=g) * 0.05, torch.randn(1, generator=g) * 0.1
    wl, bl = torch.randn(2, C, 3, 3, generator=g) * 0.05, torch.randn(2, generator=g) * 0.1
    packed = ops.pack_head_pair_tc(ws.cuda(), bs.cuda(), wl.cuda(), bl.cuda(), math="f16")
    score = torch.zeros(B, 1, H, W, device="cuda")
    shift = torch.zeros(B, 2, H, W, device="cuda")
    ops.tc_conv(_s(sh), packed, 3, src1=_s(lh), dst=score, dst_mode=3, dst_layout=1,
                dst_pool=shift).run()
    assert rel_err(score, F.conv2d(sh, ws, bs, padding=1).sigmoid()) < 2e-5
    assert rel_err(shift, F.conv2d(lh, wl, bl, padding=1).tanh()) < 2e-5
    wd, bd = torch.randn(1, C, 3, 3, generator=g) * 0.05, torch.randn(1, generator=g) * 0.1
    depth = torch.zeros(B, 1, H, W, device="cuda")
    ops.tc_conv(_s(sh), ops.pack_conv_tc(wd.cuda(), bias=bd.cuda(), math="f16"), 1, act=ops.ACT_SIGMOID,
                dst=depth, dst_layout=1, dst_c_total=1).run()
    assert rel_err(depth, F.conv2d(sh, wd, bd, padding=1).sigmoid()) < 2e-5


def test_split16_roundtrip_and_single_issuer():
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(11)
    x = (torch.randn(3, 9, 13, 40, generator=g) * torch.exp(torch.randn(3, 9, 13, 40, generator=g) * 2)).cuda()
    back = ops.unsplit16(ops.split16(x))
    # 22 significant bits, or -- below fp16's normal range for the remainder -- 2^-25 absolute
    assert bool(((back - x).abs() <= torch.maximum(x.abs() * 2.0 ** -22, torch.tensor(2.0 ** -24, device="cuda"))).all())
    # flags bit 0: one MMA-issuing thread, bit-reproducible; agrees with the three-issuer schedule to rounding
    xin = torch.randn(2, 64, 21, 45, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g) * 0.1
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="f16")
    outs = []
    for det in (True, True, False):
        o = torch.zeros(2, 64, 21, 45, device="cuda")
        ops.tc_conv(_s(xin), packed, 64, act=1, dst=o, dst_layout=1, deterministic=det).run()
        outs.append(o)
    assert torch.equal(outs[0], outs[1])
    assert rel_err(outs[2], outs[0]) < 2e-6
    assert rel_err(outs[0], _ref(xin, w, b, 1)) < 2e-5
