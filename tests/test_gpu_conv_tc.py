"""Tensor-core conv (tcgen05 3xTF32, channels-last) and the small-cout head conv vs torch fp32 on CPU."""
import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("cin,cout,H,W,act", [
    (64, 64, 60, 80, 1), (32, 32, 24, 40, 1), (64, 128, 15, 20, 0), (96, 64, 33, 50, 1), (32, 64, 17, 31, 2),
    (64, 28, 20, 28, 0), (64, 32, 9, 19, 0), (128, 64, 8, 16, 1),
    (16, 32, 40, 56, 1),   # stem layer conv1b: paired-tap steps (two taps per K = 32 row, 64-byte halo rows)
    (16, 24, 21, 37, 0),   # ... with padded output channels (N letters)
    (64, 64, 7, 5, 1),     # tile larger than the image
])
def test_conv_tc_plain_nhwc_and_nchw(cin, cout, H, W, act):
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(cin + cout + H)
    B = 3
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(x, w, b, padding=1)
    ref = F.leaky_relu(ref, 0.01) if act == 1 else (F.relu(ref) if act == 2 else ref)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    xs = _nhwc(x).cuda()
    cpad = packed[2].numel()
    if cout % 4 == 0:
        out = torch.zeros(B, H, W, cout, device="cuda")
        ops.TcConv(xs, packed, cout, act=act, dst=out, dst_layout=0).run()
        torch.cuda.synchronize()
        assert rel_err(out.permute(0, 3, 1, 2), ref) < 2e-5, rel_err(out.permute(0, 3, 1, 2), ref)
    out2 = torch.zeros(B, cout, H, W, device="cuda")
    op = ops.TcConv(xs, packed, cout, act=act, dst=None, dst_layout=1, dst_c_total=cout)
    op.run(dst_override=out2)
    torch.cuda.synchronize()
    assert rel_err(out2, ref) < 2e-5, rel_err(out2, ref)


def test_conv_tc_16_channels_single_tap_steps(monkeypatch):
    """The nine-step kernels for 16-channel inputs stay available and agree: SWIZZLE_64B single-tap steps
    (pack_conv_tc(pair_taps=False)) and paired taps (the default of round 1; the row-stationary kernel with 16-channel
    chunks serves this layer now, NVS_TC_ROW3=auto)."""
    from nano_vs_slam_b200 import ops

    monkeypatch.setenv("NVS_TC_ROW3", "32")  # 16-channel inputs stay on the nine-step kernels

    g = torch.Generator().manual_seed(16)
    B, H, W = 2, 30, 44
    x = torch.randn(B, 16, H, W, generator=g)
    w = torch.randn(32, 16, 3, 3, generator=g) * 0.1
    b = torch.randn(32, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    for pair in (False, True):
        packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), pair_taps=pair, math="tf32")
        assert packed[0].shape[0] == (5 if pair else 9)
        out = torch.zeros(B, H, W, 32, device="cuda")
        ops.TcConv(_nhwc(x).cuda(), packed, 32, act=1, dst=out).run()
        assert rel_err(out.permute(0, 3, 1, 2), ref) < 2e-5, (pair, rel_err(out.permute(0, 3, 1, 2), ref))


def test_conv_tc_pool_shuffle_concat_slice():
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 22, 38
    x = torch.randn(B, 32, H, W, generator=g)
    w = torch.randn(64, 32, 3, 3, generator=g) * 0.08
    b = torch.randn(64, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    full = torch.zeros(B, H, W, 64, device="cuda")
    pooled = torch.zeros(B, H // 2, W // 2, 64, device="cuda")
    ops.TcConv(_nhwc(x).cuda(), packed, 64, act=1, dst=full, dst_pool=pooled).run()
    assert rel_err(full.permute(0, 3, 1, 2), ref) < 2e-5
    assert rel_err(pooled.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5
    only = torch.zeros_like(pooled)
    ops.TcConv(_nhwc(x).cuda(), packed, 64, act=1, dst=None, dst_mode=0, dst_pool=only).run()
    assert rel_err(only, pooled) < 2e-6  # several MMA issuers: accumulation order (last ulp) is timing dependent
    # the single-issuer schedule is bit-reproducible
    d1, d2 = torch.zeros_like(pooled), torch.zeros_like(pooled)
    for d in (d1, d2):
        ops.TcConv(_nhwc(x).cuda(), packed, 64, act=1, dst=None, dst_mode=0, dst_pool=d, deterministic=True).run()
    assert torch.equal(d1, d2) and rel_err(d1, pooled) < 2e-6
    # pixel shuffle (odd and even sizes)
    for (hh, ww) in ((11, 19), (16, 32)):
        xs = torch.randn(B, 64, hh, ww, generator=g)
        w2 = torch.randn(128, 64, 3, 3, generator=g) * 0.05
        b2 = torch.randn(128, generator=g) * 0.1
        out = torch.zeros(B, 2 * hh, 2 * ww, 32, device="cuda")
        ops.TcConv(_nhwc(xs).cuda(), ops.pack_conv_tc(w2.cuda(), bias=b2.cuda(), math="tf32"), 128, dst=out, dst_mode=2).run()
        assert rel_err(out.permute(0, 3, 1, 2), F.pixel_shuffle(F.conv2d(xs, w2, b2, padding=1), 2)) < 2e-5
    # two sources (concat) + channel slice + BN fold
    a = torch.randn(B, 32, H, W, generator=g)
    s = torch.randn(B, 64, H, W, generator=g)
    conv = torch.nn.Conv2d(96, 64, 3, 1, 1, bias=False)
    bn = torch.nn.BatchNorm2d(64).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        ref = F.leaky_relu(bn(conv(torch.cat([a, s], 1))), 0.01)
    bnd = {k: getattr(bn, k).cuda() for k in ("weight", "bias", "running_mean", "running_var")}
    out = torch.zeros(B, H, W, 64, device="cuda")
    ops.TcConv(_nhwc(a).cuda(), ops.pack_conv_tc(conv.weight.detach().cuda(), bn=bnd, math="tf32"), 64, act=1, src1=_nhwc(s).cuda(),
               dst=out).run()
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 2e-5
    w3 = torch.randn(32, 32, 3, 3, generator=g) * 0.08
    b3 = torch.randn(32, generator=g) * 0.1
    out = torch.zeros(B, 32, H, W, device="cuda")
    ops.TcConv(_nhwc(s).cuda(), ops.pack_conv_tc(w3.cuda(), bias=b3.cuda(), math="tf32"), 32, c0_off=32, c0=32, dst=out,
               dst_layout=1).run()
    assert rel_err(out, F.conv2d(s[:, 32:], w3, b3, padding=1)) < 2e-5


def test_conv_small_and_ffma_nhwc_store():
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(8)
    B, H, W = 2, 17, 23
    x = torch.randn(B, 64, H, W, generator=g)
    for cout, act in ((1, 3), (2, 4), (3, 0)):
        w = torch.randn(cout, 64, 3, 3, generator=g) * 0.06
        b = torch.randn(cout, generator=g) * 0.1
        ref = F.conv2d(x, w, b, padding=1)
        ref = ref.sigmoid() if act == 3 else (ref.tanh() if act == 4 else ref)
        out = ops.conv_small(_nhwc(x).cuda(), ops.pack_conv_small(w.cuda(), b.cuda()), act=act)
        assert rel_err(out, ref) < 2e-5
    # FFMA conv writing channels-last (plain and pooled) for the tensor-core consumers
    w = torch.randn(32, 16, 3, 3, generator=g) * 0.1
    b = torch.randn(32, generator=g) * 0.1
    x16 = torch.randn(B, 16, 24, 36, generator=g)
    ref = F.leaky_relu(F.conv2d(x16, w, b, padding=1), 0.01)
    wp, bp = ops.pack_conv(w.cuda(), bias=b.cuda())
    full, pooled = ops.conv(x16.cuda(), wp, bp, 32, act=1, out_mode=ops.OUT_BOTH, dst_nhwc=True, dst2_nhwc=True)
    assert rel_err(full.permute(0, 3, 1, 2), ref) < 2e-5
    assert rel_err(pooled.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5


@pytest.mark.parametrize("H,W,act", [(21, 37, 1), (8, 32, 0), (40, 70, 2)])
def test_stem_conv_3_to_16_nhwc(H, W, act):
    """backbone.conv1a on the dedicated stem kernel (NCHW image -> 16-channel NHWC), incl. ragged tiles."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(H * W)
    B = 3
    x = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    w = torch.randn(16, 3, 3, 3, generator=g) * 0.3
    b = torch.randn(16, generator=g) * 0.1
    ref = F.conv2d(x, w, b, padding=1)
    ref = F.leaky_relu(ref, 0.01) if act == 1 else (F.relu(ref) if act == 2 else ref)
    wp, bp = ops.pack_conv(w.cuda(), bias=b.cuda())
    out = ops.conv(x.cuda(), wp, bp, 16, act=act, dst_nhwc=True)
    out = out[0] if isinstance(out, (tuple, list)) else out
    assert rel_err(out.permute(0, 3, 1, 2), ref) < 1e-5, rel_err(out.permute(0, 3, 1, 2), ref)
    plain = ops.conv(x.cuda(), wp, bp, 16, act=act)  # NCHW store: the generic tiled kernel
    plain = plain[0] if isinstance(plain, (tuple, list)) else plain
    assert rel_err(plain, ref) < 1e-5


def test_conv_tc_more_than_128_output_channels():
    """Layers wider than 128 channels (letter F: 256) run as one launch per 128-channel weight slice."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(21)
    B, H, W, cin, cout = 2, 18, 26, 64, 256
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    assert packed[2].numel() == 256
    xs = _nhwc(x).cuda()
    full = torch.zeros(B, H, W, cout, device="cuda")
    pooled = torch.zeros(B, H // 2, W // 2, cout, device="cuda")
    op = ops.tc_conv(xs, packed, cout, act=1, dst=full, dst_pool=pooled)
    assert isinstance(op, ops.TcConvSplit) and len(op.ops) == 2
    op.run()
    assert rel_err(full.permute(0, 3, 1, 2), ref) < 2e-5
    assert rel_err(pooled.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5
    shuf = torch.zeros(B, 2 * H, 2 * W, cout // 4, device="cuda")
    ops.tc_conv(xs, packed, cout, act=1, dst=shuf, dst_mode=2).run()
    assert rel_err(shuf.permute(0, 3, 1, 2), F.pixel_shuffle(ref, 2)) < 2e-5
    nchw = torch.zeros(B, cout, H, W, device="cuda")
    ops.tc_conv(xs, packed, cout, act=1, dst=None, dst_layout=1, dst_c_total=cout).run(dst_override=nchw)
    assert rel_err(nchw, ref) < 2e-5


@pytest.mark.parametrize("c0,c1,cout,H,W,layout", [
    (32, 0, 32, 24, 61, 0),    # three strips, the last one with a single valid column
    (32, 0, 32, 13, 30, 0),    # exactly one strip, ragged rows (13 % 4 != 0)
    (64, 0, 28, 16, 31, 1),    # segmentation output conv: 28 classes, NCHW store
    (32, 64, 32, 10, 45, 1),   # two sources (concat read), three chunks
    (96, 0, 24, 7, 90, 0),     # N letters: padded output channels
])
def test_conv_tc_row_stationary_variant(c0, c1, cout, H, W, layout, monkeypatch):
    """Layers with <= 32 output channels take the row-stationary kernel (ROW3: one pipeline step per kernel row, the
    epilogue sums three shifted partial results); it must agree with fp32 and with the 9-step kernel it replaces."""
    from nano_vs_slam_b200 import ops

    g = torch.Generator().manual_seed(c0 + c1 + cout + W)
    B, cin = 3, c0 + c1
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    s0 = _nhwc(x[:, :c0]).cuda()
    s1 = _nhwc(x[:, c0:]).cuda() if c1 else None
    outs = {}
    for flag in ("32", "16", "0"):  # 32- / 16-channel chunks / the nine-step kernel
        monkeypatch.setenv("NVS_TC_ROW3", flag)
        cpad = packed[2].numel()
        out = torch.zeros((B, H, W, cpad) if layout == 0 else (B, cout, H, W), device="cuda")
        op = ops.TcConv(s0, packed, cout, act=1, src1=s1, dst=out, dst_layout=layout)
        assert op.row3 == (flag != "0")
        op.run()
        torch.cuda.synchronize()
        outs[flag] = out[..., :cout].permute(0, 3, 1, 2) if layout == 0 else out
        assert rel_err(outs[flag], ref) < 2e-5, (flag, rel_err(outs[flag], ref))
    assert rel_err(outs["32"], outs["0"]) < 1e-5 and rel_err(outs["16"], outs["0"]) < 1e-5


@pytest.mark.parametrize("cin,H,W,mode", [(16, 40, 56, "16"), (16, 22, 62, "16"), (32, 24, 40, "32"), (32, 18, 33, "16")])
def test_conv_tc_row_stationary_max_pool(cin, H, W, mode, monkeypatch):
    """MaxPool2d(2,2) in the ROW3 epilogue (the two rows of a pooling pair live in different warps): pooled-only output
    (backbone conv1b, encoders.py:111) and full + pooled, odd and even sizes, 16-channel inputs on the 16-channel-chunk
    variant."""
    from nano_vs_slam_b200 import ops

    monkeypatch.setenv("NVS_TC_ROW3", mode)
    g = torch.Generator().manual_seed(cin + H + W)
    B = 3
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(32, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    b = torch.randn(32, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.01)
    packed = ops.pack_conv_tc(w.cuda(), bias=b.cuda(), math="tf32")
    assert packed[0].shape[0] == 9  # not the paired-tap layout
    xs = _nhwc(x).cuda()
    only = torch.zeros(B, H // 2, W // 2, 32, device="cuda")
    op = ops.TcConv(xs, packed, 32, act=1, dst=None, dst_mode=0, dst_pool=only)
    assert op.row3
    op.run()
    assert rel_err(only.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5
    full = torch.zeros(B, H, W, 32, device="cuda")
    both = torch.zeros_like(only)
    ops.TcConv(xs, packed, 32, act=1, dst=full, dst_pool=both).run()
    assert rel_err(full.permute(0, 3, 1, 2), ref) < 2e-5
    assert rel_err(both.permute(0, 3, 1, 2), F.max_pool2d(ref, 2, 2)) < 2e-5


@pytest.mark.parametrize("mode", ["32", "16"])
def test_conv_tc_row_stationary_keypoint_heads(mode, monkeypatch):
    """The fused keypoint-head conv (score | location trunks -> 3 channels, sigmoid / tanh split) on the ROW3 kernel."""
    from nano_vs_slam_b200 import ops

    monkeypatch.setenv("NVS_TC_ROW3", mode)

    g = torch.Generator().manual_seed(77)
    B, C, H, W = 2, 64, 15, 47
    sh, lh = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ws, bs = torch.randn(1, C, 3, 3, generator=g) * 0.05, torch.randn(1, generator=g) * 0.1
    wl, bl = torch.randn(2, C, 3, 3, generator=g) * 0.05, torch.randn(2, generator=g) * 0.1
    packed = ops.pack_head_pair_tc(ws.cuda(), bs.cuda(), wl.cuda(), bl.cuda(), math="tf32")
    score = torch.zeros(B, 1, H, W, device="cuda")
    shift = torch.zeros(B, 2, H, W, device="cuda")
    op = ops.TcConv(_nhwc(sh).cuda(), packed, 3, src1=_nhwc(lh).cuda(), dst=score, dst_mode=3, dst_layout=1, dst_pool=shift)
    assert op.row3
    op.run()
    assert rel_err(score, F.conv2d(sh, ws, bs, padding=1).sigmoid()) < 2e-5
    assert rel_err(shift, F.conv2d(lh, wl, bl, padding=1).tanh()) < 2e-5
