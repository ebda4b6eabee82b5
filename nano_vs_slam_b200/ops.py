"""Tensor-level wrappers over the C ABI (device pointers + current stream in, nothing else).

PyTorch is plumbing here: it owns device memory and the stream; every function hands raw pointers to
libnanovs.so.  CUDA tensors only -- there is deliberately no CPU implementation behind these calls.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import (ACT_GELU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SIGMOID_TANH, ACT_TANH,  # noqa: F401
                    IN_PLAIN, IN_S2D, IN_U8_HWC, IN_UNIT, OUT_BOTH, OUT_PLAIN, OUT_POOL, OUT_SHUFFLE, NvsConvArgs, check, lib)


# number of libnanovs kernels launched by this process (bench.py reports it as "gpu_launches")
LAUNCHES = [0]


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise _cabi.NanovsError("nanovs ops need CUDA tensors (no CPU fallback by design)")
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------
# weight packing (done once per model, see kp2dtiny._Packed)
# ----------------------------------------------------------------------------------------------
def conv_cout_tile(cout: int) -> int:
    return lib().nvs_conv_cout_tile(cout)


def conv_cin_chunk(cin: int) -> int:
    return lib().nvs_conv_cin_chunk(cin)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn: Optional[dict] = None,
              s2d: bool = False, eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """OIHW conv weight (+ optional BatchNorm running stats) -> kernel layout.

    Returns (w_packed [cin_pad][k*k][cout_pad], b_packed [cout_pad]) on the weight's device, fp32.
    BN(eval) y = (conv - mean) * gamma / sqrt(var + eps) + beta is folded as
    w' = w * s, b' = beta - mean * s with s = gamma / sqrt(var + eps)   (modules/base.py:42-43).
    ``s2d``: a 2x2 stride-2 kernel becomes a 1x1 kernel over 4*cin space-to-depth channels
    (virtual channel ci*4 + ky*2 + kx), modules/segformer.py:93-95.
    """
    w = weight.detach().to(torch.float64)
    cout, cin, kh, kw = w.shape
    if bn is not None:
        s = bn["weight"].detach().double() / torch.sqrt(bn["running_var"].detach().double() + eps)
        w = w * s.view(-1, 1, 1, 1)
        b = bn["bias"].detach().double() - bn["running_mean"].detach().double() * s
    elif bias is not None:
        b = bias.detach().double()
    else:
        b = torch.zeros(cout, dtype=torch.float64, device=w.device)
    if s2d:
        assert kh == 2 and kw == 2
        w = w.permute(1, 2, 3, 0).reshape(cin * 4, 1, cout)
        cin, taps = cin * 4, 1
    else:
        assert kh == kw and kh in (1, 3)
        taps = kh * kw
        w = w.permute(1, 2, 3, 0).reshape(cin, taps, cout)
    ck, ct = conv_cin_chunk(cin), conv_cout_tile(cout)
    cin_pad, cout_pad = _round_up(cin, ck), _round_up(cout, ct)
    wp = torch.zeros(cin_pad, taps, cout_pad, dtype=torch.float32, device=w.device)
    wp[:cin, :, :cout] = w.to(torch.float32)
    bp = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bp[:cout] = b.to(torch.float32)
    return wp.contiguous(), bp.contiguous()


# ----------------------------------------------------------------------------------------------
# conv
# ----------------------------------------------------------------------------------------------
def make_conv_args(src0: torch.Tensor, wp: torch.Tensor, bp: torch.Tensor, cout: int, *, ksize: int = 3,
                   act: int = ACT_NONE, out_mode: int = OUT_PLAIN, in_mode: int = IN_PLAIN,
                   src1: Optional[torch.Tensor] = None, dst: Optional[torch.Tensor] = None,
                   dst2: Optional[torch.Tensor] = None, c0_off: int = 0, c0: Optional[int] = None,
                   c1_off: int = 0, c1: Optional[int] = None, dst_c_off: int = 0, dst2_c_off: int = 0,
                   dst_nhwc: int = 0, dst2_nhwc: bool = False) -> NvsConvArgs:
    """Fill an NvsConvArgs for (B, C, H, W) tensors; shapes are validated here, once, at plan time.
    ``dst_nhwc``: 0 / False = NCHW, 1 / True = fp32 channels-last, 2 = the split channels-last format of the 3xFP16
    tensor-core convs (nanovs.h: NvsConvArgs.dst_nhwc)."""
    B, c0_total, inH, inW = src0.shape
    c0 = c0_total - c0_off if c0 is None else c0
    if in_mode == IN_S2D:
        H, W = inH // 2, inW // 2
    else:
        H, W = inH, inW
    a = NvsConvArgs()
    a.src0, a.weight, a.bias = src0.data_ptr(), wp.data_ptr(), bp.data_ptr()
    a.c0_total, a.c0_off, a.c0 = c0_total, c0_off, c0
    if src1 is not None:
        assert src1.shape[0] == B and tuple(src1.shape[2:]) == (inH, inW), (src0.shape, src1.shape)
        a.src1 = src1.data_ptr()
        a.c1_total, a.c1_off = src1.shape[1], c1_off
        a.c1 = src1.shape[1] - c1_off if c1 is None else c1
    else:
        a.src1, a.c1_total, a.c1_off, a.c1 = None, 0, 0, 0
    cin = (4 * c0) if in_mode == IN_S2D else (c0 + a.c1)
    assert wp.shape[0] == _round_up(cin, conv_cin_chunk(cin)), (wp.shape, cin)
    assert wp.shape[1] == ksize * ksize and wp.shape[2] == _round_up(cout, conv_cout_tile(cout))
    if out_mode in (OUT_PLAIN, OUT_BOTH) and dst_nhwc:
        assert dst is not None and tuple(dst.shape[:3]) == (B, H, W) and dst.shape[3] >= dst_c_off + cout
    elif out_mode in (OUT_PLAIN, OUT_BOTH):
        assert dst is not None and tuple(dst.shape[2:]) == (H, W) and dst.shape[0] == B
        assert dst.shape[1] >= dst_c_off + cout
    if out_mode == OUT_SHUFFLE:
        assert dst is not None and tuple(dst.shape[2:]) == (2 * H, 2 * W) and dst.shape[1] >= dst_c_off + cout // 4
    if out_mode in (OUT_POOL, OUT_BOTH) and dst2_nhwc:
        assert dst2 is not None and tuple(dst2.shape[1:3]) == (H // 2, W // 2) and dst2.shape[3] >= dst2_c_off + cout
    elif out_mode in (OUT_POOL, OUT_BOTH):
        assert dst2 is not None and tuple(dst2.shape[2:]) == (H // 2, W // 2)
        assert dst2.shape[1] >= dst2_c_off + cout
    a.dst = _ptr(dst)
    a.dst2 = _ptr(dst2)
    a.dst_c_total = (dst.shape[3] if dst_nhwc else dst.shape[1]) if dst is not None else 0
    a.dst_c_off = dst_c_off
    a.dst2_c_total = (dst2.shape[3] if dst2_nhwc else dst2.shape[1]) if dst2 is not None else 0
    a.dst2_c_off = dst2_c_off
    a.dst_nhwc, a.dst2_nhwc = int(dst_nhwc), int(dst2_nhwc)
    a.B, a.H, a.W, a.in_H, a.in_W = B, H, W, inH, inW
    a.cout, a.ksize, a.act, a.out_mode, a.in_mode = cout, ksize, act, out_mode, in_mode
    return a


def run_conv(args: NvsConvArgs) -> None:
    check(lib().nvs_conv(C.byref(args), _stream()), "nvs_conv")
    LAUNCHES[0] += 1


def conv(src0: torch.Tensor, wp: torch.Tensor, bp: torch.Tensor, cout: int, **kw):
    """Allocate outputs and run one conv (unit tests / ad-hoc use; the model uses cached plans)."""
    src0 = _req(src0)
    if kw.get("src1") is not None:
        kw["src1"] = _req(kw["src1"])
    B, _, inH, inW = src0.shape
    in_mode, out_mode = kw.get("in_mode", IN_PLAIN), kw.get("out_mode", OUT_PLAIN)
    H, W = (inH // 2, inW // 2) if in_mode == IN_S2D else (inH, inW)
    dst = dst2 = None
    if out_mode in (OUT_PLAIN, OUT_BOTH) and kw.get("dst_nhwc"):
        dst = torch.empty(B, H, W, cout, device=src0.device, dtype=torch.float32)
    elif out_mode in (OUT_PLAIN, OUT_BOTH):
        dst = torch.empty(B, cout, H, W, device=src0.device, dtype=torch.float32)
    if out_mode == OUT_SHUFFLE:
        dst = torch.empty(B, cout // 4, 2 * H, 2 * W, device=src0.device, dtype=torch.float32)
    if out_mode in (OUT_POOL, OUT_BOTH) and kw.get("dst2_nhwc"):
        dst2 = torch.empty(B, H // 2, W // 2, cout, device=src0.device, dtype=torch.float32)
    elif out_mode in (OUT_POOL, OUT_BOTH):
        dst2 = torch.empty(B, cout, H // 2, W // 2, device=src0.device, dtype=torch.float32)
    run_conv(make_conv_args(src0, wp, bp, cout, dst=dst, dst2=dst2, **kw))
    if out_mode == OUT_BOTH:
        return dst, dst2
    return dst2 if out_mode == OUT_POOL else dst


# ----------------------------------------------------------------------------------------------
# small ops
# ----------------------------------------------------------------------------------------------
def dwconv3x3(x, w, b, out=None):
    x = _req(x)
    B, Cc, H, W = x.shape
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_dwconv3x3(x.data_ptr(), w.data_ptr(), _ptr(b), out.data_ptr(), B, Cc, H, W, _stream()),
          "nvs_dwconv3x3")
    LAUNCHES[0] += 1
    return out


def channel_layernorm(x, g, b, eps: float = 1e-5, out=None):
    x = _req(x)
    B, Cc = x.shape[:2]
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_channel_layernorm(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), B, Cc,
                                      x[0, 0].numel(), eps, _stream()), "nvs_channel_layernorm")
    LAUNCHES[0] += 1
    return out


def softmax_channels(x, out=None):
    x = _req(x)
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_softmax_channels(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x[0, 0].numel(),
                                     _stream()), "nvs_softmax_channels")
    LAUNCHES[0] += 1
    return out


def l2norm_channels(x, out=None):
    x = _req(x)
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_l2norm_channels(x.data_ptr(), out.data_ptr(), x.shape[0], x.shape[1], x[0, 0].numel(),
                                    _stream()), "nvs_l2norm_channels")
    LAUNCHES[0] += 1
    return out


def attention(q, kv, heads: int, out=None):
    """q (B,C,h,w), kv (B,2C,hk,wk) -> (B,C,h,w)  (modules/segformer.py:113-133)."""
    q, kv = _req(q), _req(kv)
    B, Cc = q.shape[:2]
    out = torch.empty_like(q) if out is None else out
    check(lib().nvs_attention(q.data_ptr(), kv.data_ptr(), out.data_ptr(), B, Cc, heads, q[0, 0].numel(),
                              kv[0, 0].numel(), _stream()), "nvs_attention")
    LAUNCHES[0] += 1
    return out


def netvlad_workspace_bytes(B: int, Cc: int, K: int, S: int) -> int:
    return int(lib().nvs_netvlad_workspace_bytes(B, Cc, K, S))


def netvlad(x, w_assign, centroids, out=None, workspace=None):
    """x (B,C,h,w) -> (B, K*C)  (modules/aggregators/netvlad.py:79-106)."""
    x = _req(x)
    B, Cc = x.shape[:2]
    S = x[0, 0].numel()
    K = centroids.shape[0]
    nbytes = netvlad_workspace_bytes(B, Cc, K, S)
    if workspace is None:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty(B, K * Cc, device=x.device, dtype=torch.float32) if out is None else out
    check(lib().nvs_netvlad(x.data_ptr(), w_assign.data_ptr(), centroids.data_ptr(), out.data_ptr(),
                            workspace.data_ptr(), workspace.numel(), B, Cc, K, S, _stream()), "nvs_netvlad")
    LAUNCHES[0] += 2
    return out


def preprocess_u8(img: torch.Tensor, size=None, out=None) -> torch.Tensor:
    """uint8 HWC camera frames (B,H,W,3) [or (H,W,3)] on the device -> fp32 (B,3,H',W') in [-1,1]:
    /255, optional bilinear resize to ``size`` = (H',W') (kornia resize == F.interpolate(bilinear,
    align_corners=False)), (x - 0.5) * 2   (visual_odometry.py:281-291, frontend.py:79)."""
    if img.dim() == 3:
        img = img.unsqueeze(0)
    if not (img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3):
        raise NanovsError("preprocess_u8 expects a CUDA uint8 tensor shaped (B,H,W,3)")
    img = img.contiguous()
    B, H, W, _ = img.shape
    Ho, Wo = (H, W) if size is None else (int(size[0]), int(size[1]))
    out = torch.empty(B, 3, Ho, Wo, device=img.device, dtype=torch.float32) if out is None else out
    check(lib().nvs_preprocess_u8(img.data_ptr(), out.data_ptr(), B, H, W, Ho, Wo, _stream()), "nvs_preprocess_u8")
    LAUNCHES[0] += 1
    return out


def gem(x, p: float, eps: float = 1e-6, out=None):
    """x (B,C,h,w), h and w multiples of 4 -> (B, 16*C): GeM over PixelUnshuffle(4) (aggregators/gem.py:21-33)."""
    x = _req(x)
    B, Cc, h, w = x.shape
    out = torch.empty(B, 16 * Cc, device=x.device, dtype=torch.float32) if out is None else out
    check(lib().nvs_gem(x.data_ptr(), out.data_ptr(), B, Cc, h, w, float(p), float(eps), _stream()), "nvs_gem")
    LAUNCHES[0] += 1
    return out


def convap(x, weight, bias, s1: int = 4, s2: int = 4, out=None, workspace=None):
    """x (B,Cin,h,w) -> (B, Cout*s1*s2): 1x1 conv + AdaptiveAvgPool2d + L2 norm (aggregators/convap.py:29-37)."""
    x = _req(x)
    B, cin, h, w = x.shape
    cout = weight.shape[0]
    nbytes = int(lib().nvs_convap_workspace_bytes(B, cin, s1, s2))
    if workspace is None:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty(B, cout * s1 * s2, device=x.device, dtype=torch.float32) if out is None else out
    check(lib().nvs_convap(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), workspace.data_ptr(),
                           workspace.numel(), B, cin, cout, h, w, s1, s2, _stream()), "nvs_convap")
    LAUNCHES[0] += 2
    return out


def decode(score, shift, feat, H: int, W: int, cell: int, cross_ratio: float = 2.0):
    """post_processing core (kp2dtiny.py:593-631): returns (score_masked, coord_px, feat_sampled_unit)."""
    score, shift = _req(score), _req(shift)
    B, _, Hc, Wc = score.shape
    o_s = torch.empty_like(score)
    o_c = torch.empty_like(shift)
    if feat is not None:
        feat = _req(feat)
        D, Hf, Wf = feat.shape[1:]
        o_f = torch.empty(B, D, Hc, Wc, device=score.device, dtype=torch.float32)
    else:
        D = Hf = Wf = 0
        o_f = None
    check(lib().nvs_decode(score.data_ptr(), shift.data_ptr(), _ptr(feat), o_s.data_ptr(), o_c.data_ptr(),
                           _ptr(o_f), B, Hc, Wc, D, Hf, Wf, H, W, cell, float(cross_ratio), _stream()),
          "nvs_decode")
    LAUNCHES[0] += 1
    return o_s, o_c, o_f


def seg_argmax(seg, coord=None, H: int = 0, W: int = 0):
    """argmax over classes -> int64 (B,1,h,w); with ``coord`` nearest-sampled at keypoints first."""
    seg = _req(seg)
    B, Cc, Hs, Ws = seg.shape
    if coord is not None:
        coord = _req(coord)
        Hc, Wc = coord.shape[2:]
        out = torch.empty(B, 1, Hc, Wc, device=seg.device, dtype=torch.int64)
    else:
        Hc = Wc = 0
        out = torch.empty(B, 1, Hs, Ws, device=seg.device, dtype=torch.int64)
    check(lib().nvs_seg_argmax(seg.data_ptr(), _ptr(coord), out.data_ptr(), B, Cc, Hs, Ws, Hc, Wc, H, W,
                               _stream()), "nvs_seg_argmax")
    LAUNCHES[0] += 1
    return out


def select_keypoints(score, coord, feat, thresh: float, top_k: int, seg_cells=None,
                     classes_to_filter: Optional[Sequence[int]] = None):
    """Device version of frontend.py:94-126 for a whole batch.

    score (B,1,Hc,Wc), coord (B,2,Hc,Wc), feat (B,D,Hc,Wc) [, seg_cells (B,1,Hc,Wc) int64]
    -> dict(pts (B,k,2), desc (B,k,D), score (B,k), cell (B,k) int32, label (B,k) int64|None, count (B,) int32)
    Rows >= count[b] are unspecified.
    """
    score, coord = _req(score), _req(coord)
    B = score.shape[0]
    n_cells = score[0].numel()
    k = n_cells if top_k <= 0 else min(top_k, n_cells)
    dev = score.device
    D = 0
    if feat is not None:
        feat = _req(feat)
        D = feat.shape[1]
    pts = torch.empty(B, k, 2, device=dev, dtype=torch.float32)
    desc = torch.empty(B, k, D, device=dev, dtype=torch.float32) if feat is not None else None
    sc = torch.empty(B, k, device=dev, dtype=torch.float32)
    cell = torch.empty(B, k, device=dev, dtype=torch.int32)
    count = torch.empty(B, device=dev, dtype=torch.int32)
    label = filt = None
    if seg_cells is not None:
        seg_cells = _req(seg_cells, torch.int64)
        assert seg_cells[0].numel() == n_cells, "seg labels must be per cell (sample_segmentation=True)"
        label = torch.empty(B, k, device=dev, dtype=torch.int64)
        if classes_to_filter:
            filt = torch.tensor(list(classes_to_filter), device=dev, dtype=torch.int32)
    check(lib().nvs_select_keypoints(score.data_ptr(), coord.data_ptr(), _ptr(feat), _ptr(seg_cells), _ptr(filt),
                                     0 if filt is None else filt.numel(), float(thresh), k, pts.data_ptr(),
                                     _ptr(desc), sc.data_ptr(), cell.data_ptr(), _ptr(label), count.data_ptr(),
                                     B, n_cells, D, _stream()), "nvs_select_keypoints")
    LAUNCHES[0] += 1
    return {"pts": pts, "desc": desc, "score": sc, "cell": cell, "label": label, "count": count}


def match(des1, des2, ratio: float = 0.7, mode: int = 0):
    """mode 0: 2-NN + ratio + one-to-one (feature_matcher.py:89-98,179-209); 1: mutual NN; 2: raw 2-NN.

    Returns device tensors: modes 0/1 -> (idx1, idx2, dist, count[1]); mode 2 -> (idx (n1,2), dist (n1,2)).
    """
    des1, des2 = _req(des1), _req(des2)
    n1, D = des1.shape
    n2 = des2.shape[0]
    dev = des1.device
    ws = torch.empty(int(lib().nvs_match_workspace_bytes(n1, n2)), dtype=torch.uint8, device=dev)
    if mode == 2:
        idx = torch.empty(n1, 2, device=dev, dtype=torch.int32)
        dist = torch.empty(n1, 2, device=dev, dtype=torch.float32)
        dummy = torch.empty(1, device=dev, dtype=torch.int32)
        check(lib().nvs_match(des1.data_ptr(), des2.data_ptr(), n1, n2, D, float(ratio), 2, idx.data_ptr(),
                              dummy.data_ptr(), dist.data_ptr(), dummy.data_ptr(), ws.data_ptr(), ws.numel(),
                              _stream()), "nvs_match")
        LAUNCHES[0] += 2
        return idx, dist
    i1 = torch.empty(n1, device=dev, dtype=torch.int32)
    i2 = torch.empty(n1, device=dev, dtype=torch.int32)
    dd = torch.empty(n1, device=dev, dtype=torch.float32)
    cnt = torch.zeros(1, device=dev, dtype=torch.int32)
    check(lib().nvs_match(des1.data_ptr(), des2.data_ptr(), n1, n2, D, float(ratio), mode, i1.data_ptr(),
                          i2.data_ptr(), dd.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "nvs_match")
    LAUNCHES[0] += 3 if mode == 0 else 5
    return i1, i2, dd, cnt


def match_batch(desc: torch.Tensor, counts: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor,
                ratio: float = 0.7, mode: int = 0, workspace: Optional[torch.Tensor] = None):
    """All pairs in one call, nothing returns to the host: ``desc`` (F,kmax,D) and ``counts`` (F,) int32 as written
    by select_keypoints, ``pair_a`` / ``pair_b`` (P,) int32 frame indices on the device.
    Returns (idx1 (P,kmax), idx2 (P,kmax), dist (P,kmax), count (P,)); rows >= count[p] are unspecified."""
    desc = _req(desc)
    F_, kmax, D = desc.shape
    dev = desc.device
    counts = counts.to(device=dev, dtype=torch.int32).contiguous()
    pair_a = pair_a.to(device=dev, dtype=torch.int32).contiguous()
    pair_b = pair_b.to(device=dev, dtype=torch.int32).contiguous()
    P = pair_a.numel()
    nbytes = int(lib().nvs_match_batch_workspace_bytes(P, kmax))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    i1 = torch.empty(P, kmax, device=dev, dtype=torch.int32)
    i2 = torch.empty(P, kmax, device=dev, dtype=torch.int32)
    dd = torch.empty(P, kmax, device=dev, dtype=torch.float32)
    cnt = torch.empty(P, device=dev, dtype=torch.int32)
    check(lib().nvs_match_batch(desc.data_ptr(), counts.data_ptr(), F_, kmax, D, pair_a.data_ptr(), pair_b.data_ptr(), P,
                                float(ratio), mode, i1.data_ptr(), i2.data_ptr(), dd.data_ptr(), cnt.data_ptr(),
                                workspace.data_ptr(), workspace.numel(), _stream()), "nvs_match_batch")
    LAUNCHES[0] += 3 if mode == 0 else 5
    return i1, i2, dd, cnt


def pose_batch(pts: torch.Tensor, pair_a: torch.Tensor, pair_b: torch.Tensor, count: torch.Tensor,
               idx1: Optional[torch.Tensor] = None, idx2: Optional[torch.Tensor] = None,
               intrinsics: Sequence[float] = (1.0, 1.0, 0.0, 0.0), threshold: float = 0.0003, iters: int = 512,
               seed: int = 0, refine: int = 0, workspace: Optional[torch.Tensor] = None, confidence: float = 0.0,
               round_size: int = 64):
    """Relative pose of P frame pairs in one call (visual_odometry.py:383-412: findEssentialMat + recoverPose).

    ``confidence`` > 0 (the reference passes prob = 0.999): data-dependent sample count -- samples are drawn in rounds of
    ``round_size`` and a pair stops once its count reaches the RANSAC bound for its best consensus so far, at most
    ``iters`` samples; the result then carries "iters" (P,) int32, the samples evaluated per pair.  0 (default): exactly
    ``iters`` samples per pair (fixed cost; what bench.py times).

    ``pts`` (F,kmax,2) keypoint coordinates as written by select_keypoints, ``pair_a`` (current) / ``pair_b``
    (reference) (P,) frame indices, ``idx1`` / ``idx2`` / ``count`` as returned by match_batch (both idx None: the
    rows of ``pts`` are already matched).  ``intrinsics`` = (fx, fy, cx, cy) of the pinhole camera.
    ``refine`` = 0 returns the best minimal-sample model (cv2.RANSAC behaviour); n > 0 adds up to n Gauss-Newton steps on
    the consensus set (the local optimisation / polishing of cv2.USAC_MSAC).
    Returns dict(E (P,3,3), R (P,3,3), t (P,3), mask (P,kmax) uint8, inliers (P,) int32): x_ref ~ R x_cur + t."""
    pts = _req(pts)
    F_, kmax, two = pts.shape
    assert two == 2
    dev = pts.device
    pair_a = pair_a.to(device=dev, dtype=torch.int32).contiguous()
    pair_b = pair_b.to(device=dev, dtype=torch.int32).contiguous()
    count = count.to(device=dev, dtype=torch.int32).contiguous()
    P = pair_a.numel()
    if (idx1 is None) != (idx2 is None):
        raise ValueError("idx1 and idx2 go together")
    if idx1 is not None:
        idx1, idx2 = _req(idx1, torch.int32), _req(idx2, torch.int32)
        assert idx1.shape == (P, kmax) and idx2.shape == (P, kmax)
    nbytes = int(lib().nvs_pose_workspace_bytes(P, kmax, iters))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    E = torch.empty(P, 3, 3, device=dev, dtype=torch.float32)
    R = torch.empty(P, 3, 3, device=dev, dtype=torch.float32)
    t = torch.empty(P, 3, device=dev, dtype=torch.float32)
    mask = torch.empty(P, kmax, device=dev, dtype=torch.uint8)
    inl = torch.empty(P, device=dev, dtype=torch.int32)
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    if confidence > 0.0:
        used = torch.empty(P, device=dev, dtype=torch.int32)
        check(lib().nvs_pose_batch_adaptive(pts.data_ptr(), F_, kmax, pair_a.data_ptr(), pair_b.data_ptr(), _ptr(idx1),
                                            _ptr(idx2), count.data_ptr(), P, fx, fy, cx, cy, float(threshold), int(iters),
                                            int(seed) & 0xFFFFFFFFFFFFFFFF, int(refine), float(confidence),
                                            int(round_size), E.data_ptr(), R.data_ptr(), t.data_ptr(), mask.data_ptr(),
                                            inl.data_ptr(), used.data_ptr(), workspace.data_ptr(), workspace.numel(),
                                            _stream()), "nvs_pose_batch_adaptive")
        LAUNCHES[0] += 2 + 3 * ((iters + round_size - 1) // round_size) + (1 if refine > 0 else 0)
        return {"E": E, "R": R, "t": t, "mask": mask, "inliers": inl, "iters": used}
    check(lib().nvs_pose_batch(pts.data_ptr(), F_, kmax, pair_a.data_ptr(), pair_b.data_ptr(), _ptr(idx1), _ptr(idx2),
                               count.data_ptr(), P, fx, fy, cx, cy, float(threshold), int(iters),
                               int(seed) & 0xFFFFFFFFFFFFFFFF, int(refine), E.data_ptr(), R.data_ptr(), t.data_ptr(),
                               mask.data_ptr(), inl.data_ptr(), workspace.data_ptr(), workspace.numel(), _stream()),
          "nvs_pose_batch")
    LAUNCHES[0] += 5 if refine > 0 else 4
    return {"E": E, "R": R, "t": t, "mask": mask, "inliers": inl}


# ----------------------------------------------------------------------------------------------
# tensor-core conv (csrc/conv_tc.cu): channels-last activations, 3xTF32 weights
# ----------------------------------------------------------------------------------------------
def conv_tc_supported(c0: int, c1: int, cout: int) -> bool:
    return bool(lib().nvs_conv_tc_supported(c0, c1, cout))


def _fold(weight, bias, bn, eps):
    w = weight.detach().to(torch.float64)
    cout = w.shape[0]
    if bn is not None:
        s = bn["weight"].detach().double() / torch.sqrt(bn["running_var"].detach().double() + eps)
        w = w * s.view(-1, 1, 1, 1)
        b = bn["bias"].detach().double() - bn["running_mean"].detach().double() * s
    elif bias is not None:
        b = bias.detach().double()
    else:
        b = torch.zeros(cout, dtype=torch.float64, device=w.device)
    return w.to(torch.float32), b.to(torch.float32)


def row3_mode() -> str:
    """Which variant of the row-stationary tensor-core conv serves layers with <= 32 output channels (env
    NVS_TC_ROW3): "0" = none (nine single-tap pipeline steps), "auto" (default, also "1") = 16-channel chunks for a
    16-channel input (the stem's second conv), 32-channel chunks otherwise, "32" = 32-channel chunks only (16-channel
    inputs keep the paired-tap kernel), "16" = 16-channel chunks everywhere."""
    import os
    v = os.environ.get("NVS_TC_ROW3", "auto")
    return {"0": "0", "1": "auto", "auto": "auto", "32": "32", "16": "16"}.get(v, "auto")


def conv_math() -> str:
    """Arithmetic of the tensor-core 3x3 convs (env NVS_CONV_MATH): "f16" (default) = the row-stationary "3xFP16" kernel
    of csrc/conv_rs.cu (fp16 hi / lo operand pairs, K = 16 per MMA, operands from shared memory); "tf32" = the 3xTF32
    kernels of csrc/conv_tc.cu (TMEM-fed, 8-bit exponent: no restriction on the activation range)."""
    import os
    v = os.environ.get("NVS_CONV_MATH", "f16").lower()
    return "tf32" if v in ("tf32", "3xtf32") else "f16"


class RsPacked(object):
    """Weights of one 3x3 conv for the "3xFP16" kernel: per slice of <= 64 output channels the fp16 hi / lo parts of
    w * 2^t in the layout [3 ky][3 kx x cout_pad][cin] (cout_pad = 32 or 64), the fp32 bias [cout_pad] and 2^-t."""

    def __init__(self, slices, cout: int, cin: int, segments):
        self.slices = slices          # [(hi16, lo16, bias32, w_scale, cout_slice)]
        self.cout, self.cin = cout, cin
        self.nvs_segments = segments  # [(real, padded)] per source, or None


def pack_conv_rs(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn: Optional[dict] = None,
                 eps: float = 1e-5, cin_segments=None) -> RsPacked:
    """OIHW 3x3 weight (+BN) -> RsPacked for nvs_conv_tc with flags bit 4 (csrc/conv_rs.cu).

    The folded fp32 weights are scaled by 2^t (one exponent per layer, exact) so that the largest magnitude lands in
    [2^13, 2^14): hi = fp16(w 2^t), lo = fp16(w 2^t - hi) -- both normal fp16 numbers for every weight down to
    2^-27 of the largest; the kernel multiplies its accumulator by 2^-t.  Input-channel segments as in pack_conv_tc; a
    16-channel input (the stem's output) is padded to one 32-channel chunk."""
    w, b = _fold(weight, bias, bn, eps)
    if cin_segments is not None:
        assert sum(r for r, _ in cin_segments) == w.shape[1], (cin_segments, w.shape)
        parts, at = [], 0
        for real, padded in cin_segments:
            seg = w[:, at:at + real]
            if padded > real:
                seg = torch.cat([seg, torch.zeros(w.shape[0], padded - real, 3, 3, dtype=w.dtype, device=w.device)], 1)
            parts.append(seg)
            at += real
        w = torch.cat(parts, 1)
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    if cin == 16:
        w = torch.cat([w, torch.zeros(cout, 16, 3, 3, dtype=w.dtype, device=w.device)], 1)
        cin = 32
    assert cin % 32 == 0, cin
    amax = float(w.abs().max())
    t = 0 if amax == 0.0 else int(torch.floor(torch.log2(torch.tensor(16383.0 / amax))).item())
    t = max(-14, min(t, 40))
    ws = w.double() * (2.0 ** t)
    slices = []
    for j in range(0, cout, 64):
        cj = min(64, cout - j)
        cpad = 32 if cj <= 32 else 64
        wt = torch.zeros(3, 3, cpad, cin, dtype=torch.float64, device=w.device)     # [ky][kx][co][ci]
        wt[:, :, :cj] = ws[j:j + cj].permute(2, 3, 0, 1)
        wt = wt.reshape(3, 3 * cpad, cin)
        hi = wt.to(torch.float16)
        lo = (wt - hi.double()).to(torch.float16)
        bp = torch.zeros(cpad, dtype=torch.float32, device=w.device)
        bp[:cj] = b[j:j + cj]
        slices.append((hi.contiguous(), lo.contiguous(), bp, float(2.0 ** -t), cj))
    return RsPacked(slices, cout, cin, list(cin_segments) if cin_segments is not None else None)


def pack_conv_tc(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, bn: Optional[dict] = None,
                 eps: float = 1e-5, cin_segments=None, pair_taps: Optional[bool] = None, math: Optional[str] = None):
    """OIHW 3x3 weight (+BN) -> (w_hi, w_lo) [9][cout_pad][cin] and bias [cout_pad] for nvs_conv_tc -- or, with
    ``math`` "f16" (the default, see conv_math), the RsPacked weights of the "3xFP16" kernel.

    w_hi = w rounded to tf32 (10 explicit mantissa bits), w_lo = tf32-rounded (w - w_hi).
    ``cin_segments``: [(real, padded), ...] -- the input channels arrive as consecutive segments of ``real``
    channels, each stored in a channels-last buffer padded with zeros to ``padded`` channels (the N letters have
    24/48/72/96 channels; the tensor-core kernel works on 32-channel rows).  Zero weight columns are inserted."""
    if math is None:
        math = conv_math()
    if math == "f16" and pair_taps is None:
        return pack_conv_rs(weight, bias=bias, bn=bn, eps=eps, cin_segments=cin_segments)
    w, b = _fold(weight, bias, bn, eps)
    if cin_segments is not None:
        assert sum(r for r, _ in cin_segments) == w.shape[1], (cin_segments, w.shape)
        parts, at = [], 0
        for real, padded in cin_segments:
            seg = w[:, at:at + real]
            if padded > real:
                seg = torch.cat([seg, torch.zeros(w.shape[0], padded - real, 3, 3, dtype=w.dtype, device=w.device)], 1)
            parts.append(seg)
            at += real
        w = torch.cat(parts, 1)
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    # more than 128 output channels: padded to a multiple of 128 and run as several launches (TcConvSplit)
    cpad = int(lib().nvs_conv_tc_cout_pad(cout)) if cout <= 128 else (cout + 127) // 128 * 128
    assert cpad > 0, cout
    wt = torch.zeros(9, cpad, cin, dtype=torch.float32, device=w.device)
    wt[:, :cout] = w.permute(2, 3, 0, 1).reshape(9, cout, cin)
    if pair_taps is None:  # 16-channel inputs: paired taps unless the 16-channel row-stationary kernel takes the layer
        pair_taps = row3_mode() in ("0", "32")
    if cin == 16 and cpad == 32 and pair_taps:
        # paired-tap layout for 16-channel inputs (NvsConvTcArgs.flags bit 1): K row of step t =
        # [tap 2t, channels 0-15 | tap 2t+1, channels 0-15]; the tenth tap is zero
        wp = torch.zeros(10, cpad, 16, dtype=torch.float32, device=w.device)
        wp[:9] = wt
        wt = wp.view(5, 2, cpad, 16).permute(0, 2, 1, 3).reshape(5, cpad, 32).contiguous()
    # hi = w rounded to the nearest tf32 value, lo = (w - hi) rounded to the nearest tf32 value (the tensor core
    # truncates operands to tf32; pre-rounding makes the residual error unbiased)
    hi = ((wt.view(torch.int32) + 0x1000) & -8192).view(torch.float32).contiguous()
    lo = (wt - hi)
    lo = ((lo.view(torch.int32) + 0x1000) & -8192).view(torch.float32).contiguous()
    bp = torch.zeros(cpad, dtype=torch.float32, device=w.device)
    bp[:cout] = b
    # (real, padded) channels per source: lets the kernel skip MMA k-steps that only see zero weights
    hi.nvs_segments = list(cin_segments) if cin_segments is not None else None
    return hi, lo, bp


def pack_head_pair_tc(w_score: torch.Tensor, b_score: torch.Tensor, w_shift: Optional[torch.Tensor] = None,
                      b_shift: Optional[torch.Tensor] = None, cpad: Optional[int] = None, math: Optional[str] = None):
    """Keypoint-head output conv(s) as ONE tensor-core conv with 3 output channels.

    V2: score_head.convDb (1, C, 3, 3) reads the score trunk and loc_head.convDb (2, C, 3, 3) reads the location
    trunk (heads.py:33): the two trunks are the two sources of the conv, the weight is block diagonal (3, 2C, 3, 3).
    V3: score_loc_head.convDb (3, C, 3, 3) is used as is (pass it as ``w_score``)."""
    c = w_score.shape[1]
    cpad = c if cpad is None else cpad
    if w_shift is None:
        return pack_conv_tc(w_score, bias=b_score, cin_segments=[(c, cpad)], math=math)
    w = torch.zeros(3, 2 * c, 3, 3, dtype=torch.float32, device=w_score.device)
    w[0:1, :c] = w_score.detach().float()
    w[1:3, c:] = w_shift.detach().float()
    return pack_conv_tc(w, bias=torch.cat([b_score.detach().float(), b_shift.detach().float()]),
                        cin_segments=[(c, cpad), (c, cpad)], math=math)


def pack_conv_small(weight: torch.Tensor, bias: torch.Tensor):
    """OIHW (cout<=4) -> ([9][cout][cin], bias[cout]) for nvs_conv_small."""
    w = weight.detach().float()
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous(), bias.detach().float().contiguous()


class TcConv(object):
    """One planned nvs_conv_tc launch: owns the plan memory (TMA descriptors) and keeps operands alive.

    All activation tensors are channels-last *shaped* torch tensors: (B, H, W, C)."""

    def __init__(self, src0: torch.Tensor, packed, cout: int, *, act: int = ACT_NONE,
                 src1: Optional[torch.Tensor] = None, c0_off: int = 0, c0: Optional[int] = None, c1_off: int = 0,
                 c1: Optional[int] = None, dst: Optional[torch.Tensor] = None, dst_layout: int = 0,
                 dst_mode: int = 1, dst_c_off: int = 0, dst_c_total: Optional[int] = None,
                 dst_pool: Optional[torch.Tensor] = None, pool_c_off: int = 0,
                 deterministic: Optional[bool] = None, rs_scale: Optional[float] = None, rs_segments=None):
        """``deterministic``: one MMA-issuing thread instead of two -> fixed fp32 accumulation order, results
        bit-reproducible run to run (default: env NVS_DETERMINISTIC=1, else the faster multi-issuer schedule whose
        last-ulp rounding depends on timing)."""
        import os
        hi, lo, bp = packed
        if deterministic is None:
            deterministic = os.environ.get("NVS_DETERMINISTIC", "0") == "1"
        B, H, W, c0_total = src0.shape
        a = _cabi.NvsConvTcArgs()
        a.src0, a.w_hi, a.w_lo, a.bias = src0.data_ptr(), hi.data_ptr(), lo.data_ptr(), bp.data_ptr()
        a.c0_total, a.c0_off = c0_total, c0_off
        a.c0 = c0_total - c0_off if c0 is None else c0
        if src1 is not None:
            assert tuple(src1.shape[:3]) == (B, H, W)
            a.src1, a.c1_total, a.c1_off = src1.data_ptr(), src1.shape[3], c1_off
            a.c1 = src1.shape[3] - c1_off if c1 is None else c1
        else:
            a.src1, a.c1_total, a.c1_off, a.c1 = None, 0, 0, 0
        rs = rs_scale is not None  # one slice of an RsPacked: the "3xFP16" row-stationary kernel (flags bit 4)
        paired = (not rs) and hi.shape[0] == 5  # pack_conv_tc's paired-tap layout for 16-channel inputs
        if rs:
            assert hi.dtype == torch.float16 and hi.shape[0] == 3 and hi.shape[1] == 3 * bp.numel(), hi.shape
            assert hi.shape[2] == (32 if (a.c0 == 16 and a.c1 == 0) else a.c0 + a.c1), (hi.shape, a.c0, a.c1)
        else:
            assert (hi.shape[2] == 32 and a.c0 == 16 and a.c1 == 0) if paired else hi.shape[2] == a.c0 + a.c1, \
                (hi.shape, a.c0, a.c1)
        a.dst = _ptr(dst)
        a.dst_pool = _ptr(dst_pool)
        if dst_c_total is None:
            if dst is None:
                dst_c_total = cout // 4 if dst_mode == 2 else cout
            else:
                dst_c_total = dst.shape[1] if dst_layout == 1 else dst.shape[3]
        a.dst_c_total, a.dst_c_off, a.dst_layout, a.dst_mode = dst_c_total, dst_c_off, dst_layout, dst_mode
        a.pool_c_total = dst_pool.shape[3] if (dst_pool is not None and dst_mode != 3) else 0
        a.pool_c_off = pool_c_off
        a.B, a.H, a.W, a.cout, a.act = B, H, W, cout, act
        # row-stationary kernel (conv_tc.cu Cfg: ROW3) for layers with <= 32 output channels: the packed [9][32][cin]
        # weights ARE its [3 ky][3 kx x 32][cin] layout, so only the flag differs (chunk width: see row3_mode).
        mode = row3_mode()
        chunk = 16 if (mode == "16" or (mode == "auto" and a.c0 == 16 and a.c1 == 0)) else 32
        row3 = (mode != "0" and bp.numel() == 32 and not paired and a.c0 % chunk == 0 and a.c1 % chunk == 0
                and dst_mode != 2 and (dst_mode != 0 or dst_pool is not None) and not rs)
        self.row3 = row3
        a.flags = ((1 if deterministic else 0) | (2 if paired else 0) | (4 if row3 else 0) |
                   (8 if row3 and chunk == 16 else 0) | (16 if rs else 0))
        a.w_scale = float(rs_scale) if rs else 1.0
        segs = rs_segments if rs else getattr(hi, "nvs_segments", None)
        a.c0_real = a.c1_real = 0
        if segs is not None and len(segs) == (2 if src1 is not None else 1) and not paired:
            if segs[0][1] == a.c0 and (src1 is None or segs[1][1] == a.c1):
                a.c0_real = segs[0][0]
                a.c1_real = segs[1][0] if src1 is not None else 0
        if dst is not None:
            if dst_mode == 1 and dst_layout == 0:
                assert tuple(dst.shape[:3]) == (B, H, W)
            elif dst_mode == 1:
                assert dst.shape[0] == B and tuple(dst.shape[2:]) == (H, W)
            elif dst_mode == 2:
                assert tuple(dst.shape[:3]) == (B, 2 * H, 2 * W)
        if dst_mode == 3:
            assert cout == 3 and act == ACT_NONE
        if dst_pool is not None and dst_mode != 3:
            assert tuple(dst_pool.shape[:3]) == (B, H // 2, W // 2)
        self._keep = (src0, src1, hi, lo, bp, dst, dst_pool)
        self._mem = C.create_string_buffer(int(lib().nvs_conv_tc_plan_bytes()))
        check(lib().nvs_conv_tc_plan_init(self._mem, C.byref(a)), "nvs_conv_tc_plan_init")
        self.flops = 2.0 * 9 * (a.c0 + a.c1) * cout * H * W * B
        self.cout_ = cout
        self.shape = (f"{a.c0 + a.c1}->{cout} k3 @{H}x{W} tcgen05" + (f" row3/{chunk}" if row3 else "") +
                      (" rs/f16" if rs else ""))

    def run(self, dst_override: Optional[torch.Tensor] = None, dst2_override: Optional[torch.Tensor] = None) -> None:
        check(lib().nvs_conv_tc_run(self._mem, _ptr(dst_override), _ptr(dst2_override), _stream()),
              "nvs_conv_tc_run")
        LAUNCHES[0] += 1


class TcConvSplit(object):
    """A tensor-core conv with more than 128 output channels (letters D / F: 256- and 512-channel layers) as one
    TcConv per 128-channel slice of the weights: each launch writes its channel range of the shared outputs."""

    def __init__(self, src0: torch.Tensor, packed, cout: int, slice_width: int = 128, **kw):
        hi, lo, bp = packed
        assert hi.shape[0] == 9, "paired-tap packing is for 32-channel outputs only"
        mode = kw.get("dst_mode", 1)
        assert mode in (0, 1, 2), "the keypoint-head split epilogue has 3 output channels"
        if kw.get("dst") is None and kw.get("dst_c_total") is None and mode != 0:
            kw["dst_c_total"] = cout // 4 if mode == 2 else cout
        import os
        # slice width: 128 (fewest launches) or 64 (the 64-wide kernel keeps the small correction products in their
        # own accumulator half: lower rounding error)
        sw = int(os.environ.get("NVS_TC_SPLIT", str(slice_width)))
        assert sw in (64, 128)
        self.ops = []
        for j in range((cout + sw - 1) // sw):
            cj = min(sw, cout - sw * j)
            hj = hi[:, sw * j:sw * (j + 1)].contiguous()
            hj.nvs_segments = getattr(hi, "nvs_segments", None)
            sub = (hj, lo[:, sw * j:sw * (j + 1)].contiguous(), bp[sw * j:sw * (j + 1)].contiguous())
            kj = dict(kw)
            kj["dst_c_off"] = kw.get("dst_c_off", 0) + (sw // 4 if mode == 2 else sw) * j
            if kw.get("dst_pool") is not None:
                kj["pool_c_off"] = kw.get("pool_c_off", 0) + sw * j
            self.ops.append(TcConv(src0, sub, cj, **kj))
        self.flops = sum(o.flops for o in self.ops)
        self.shape = self.ops[0].shape.replace(f"->{min(sw, cout)} ", f"->{cout} ") + f" x{len(self.ops)}"

    def run(self, dst_override: Optional[torch.Tensor] = None, dst2_override: Optional[torch.Tensor] = None) -> None:
        for o in self.ops:
            o.run(dst_override, dst2_override)


def split16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 channels-last (..., C) -> the split format the 3xFP16 convs read and write: same shape and dtype as a
    container, per pixel C fp16 values a_hi = fp16(a) followed by C fp16 values a_lo = fp16(a - a_hi)."""
    x = _req(x)
    assert x.shape[-1] % 8 == 0, x.shape
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_split16(x.data_ptr(), out.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], _stream()), "nvs_split16")
    LAUNCHES[0] += 1
    return out


def unsplit16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Split channels-last format -> fp32 (a_hi + a_lo)."""
    x = _req(x)
    assert x.shape[-1] % 8 == 0, x.shape
    out = torch.empty_like(x) if out is None else out
    check(lib().nvs_unsplit16(x.data_ptr(), out.data_ptr(), x.numel() // x.shape[-1], x.shape[-1], _stream()), "nvs_unsplit16")
    LAUNCHES[0] += 1
    return out


def conv_rs_range_flag(reset: bool = False) -> int:
    """1 if a "3xFP16" conv on the current device wrote an activation beyond the fp16 range since the last reset (the
    following layer's operands were then not finite: switch that model to NVS_CONV_MATH=tf32).  Synchronises."""
    v = int(lib().nvs_conv_rs_range_flag(1 if reset else 0))
    if v < 0:
        raise NanovsError("nvs_conv_rs_range_flag failed")
    return v


class RsConv(object):
    """A 3x3 conv on the "3xFP16" row-stationary kernel: one launch per <= 64-channel slice of an RsPacked, each writing
    its channel range of the shared outputs (same contract as TcConv / TcConvSplit)."""

    def __init__(self, src0: torch.Tensor, packed: RsPacked, cout: int, **kw):
        mode = kw.get("dst_mode", 1)
        n = len(packed.slices)
        assert mode != 3 or n == 1, "the keypoint-head split epilogue has 3 output channels"
        if n > 1 and kw.get("dst") is None and kw.get("dst_c_total") is None and mode != 0:
            kw["dst_c_total"] = cout // 4 if mode == 2 else cout
        self.ops = []
        covered = 0
        for j, (hi, lo, bp, scale, _real) in enumerate(packed.slices):
            # ``cout`` may count zero-weight padding channels (N letters: 24 real channels in a 32-channel buffer)
            cj = min(bp.numel(), cout - 64 * j)
            if cj <= 0:
                break
            kj = dict(kw)
            kj["dst_c_off"] = kw.get("dst_c_off", 0) + (16 if mode == 2 else 64) * j
            if kw.get("dst_pool") is not None and mode != 3:
                kj["pool_c_off"] = kw.get("pool_c_off", 0) + 64 * j
            self.ops.append(TcConv(src0, (hi, lo, bp), cj, rs_scale=scale, rs_segments=packed.nvs_segments, **kj))
            covered = 64 * j + cj
        if covered < cout:
            # the caller's (padded) channel count exceeds the packed slices (N letters: 4 x 24 real channels requested
            # as 4 x 32): the remaining channels are zero weights + zero bias = exact zeros, written once here into the
            # persistent plan buffers instead of by a launch
            assert mode in (1, 2) or kw.get("dst_pool") is not None
            dst, pool = kw.get("dst"), kw.get("dst_pool")
            assert (dst is not None or mode == 0) and kw.get("dst_layout", 0) == 0, "padding channels need a plan buffer"
            if dst is not None:
                off = kw.get("dst_c_off", 0)
                if mode == 2:
                    dst[..., off + covered // 4: off + cout // 4].zero_()
                else:
                    dst[..., off + covered: off + cout].zero_()
            if pool is not None:
                off = kw.get("pool_c_off", 0)
                pool[..., off + covered: off + cout].zero_()
        self.flops = sum(o.flops for o in self.ops)
        self.shape = self.ops[0].shape.replace(f"->{self.ops[0].cout_} ", f"->{cout} ") + (f" x{len(self.ops)}" if n > 1 else "")

    def run(self, dst_override: Optional[torch.Tensor] = None, dst2_override: Optional[torch.Tensor] = None) -> None:
        for o in self.ops:
            o.run(dst_override, dst2_override)


def tc_conv(src0: torch.Tensor, packed, cout: int, slice_width: int = 128, **kw):
    """TcConv, or TcConvSplit when the packed weights carry more than ``slice_width`` (padded) output channels; RsConv
    for RsPacked weights (NVS_CONV_MATH=f16, the default)."""
    if isinstance(packed, RsPacked):
        return RsConv(src0, packed, cout, **kw)
    if packed[2].numel() > slice_width and packed[0].shape[0] == 9 and kw.get("dst_mode", 1) != 3:
        return TcConvSplit(src0, packed, cout, slice_width=slice_width, **kw)
    return TcConv(src0, packed, cout, **kw)


def conv_small(src_nhwc: torch.Tensor, packed, act: int = ACT_NONE, out: Optional[torch.Tensor] = None):
    """(B,H,W,cin) channels-last -> (B,cout,H,W), cout <= 4 (score / location heads)."""
    w, b = packed
    B, H, W, cin = src_nhwc.shape
    cout = w.shape[1]
    if out is None:
        out = torch.empty(B, cout, H, W, device=src_nhwc.device, dtype=torch.float32)
    check(lib().nvs_conv_small(src_nhwc.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, cin, cout,
                               act, _stream()), "nvs_conv_small")
    LAUNCHES[0] += 1
    return out
