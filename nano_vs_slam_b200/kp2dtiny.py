"""KP2DTinyV2 / KP2DTinyV3 with the reference's constructor, config letters, state_dict names and
forward / post_processing output dict, executed by the sm_100a kernels in libnanovs.so.

Mirrors src/kp2dtiny/models/kp2dtiny.py of ETH-PBL/Nano-VS-SLAM:
  KP2DTINY_CONFIGS :198, KP2DTINYV3_CONFIGS :210, tiny_factory :221, get_config :245,
  KP2DTinyV2 :284 (forward :552, post_processing :593), KP2DTinyV3 :650 (forward :906,
  post_processing :959).

The nn.Module tree below only *holds parameters* under the reference's names (so reference ``.ckpt``
state_dicts load with strict=True); it never calls ATen convolutions.  ``forward`` folds BatchNorm into
packed weights once, builds a per-(B,H,W) launch plan (pre-filled C-ABI argument structs + cached
intermediate buffers) and replays it.  Inference only: the module raises in training mode and on CPU
tensors -- there is no fallback path.
"""
from __future__ import annotations

import copy
import inspect
import os
from typing import Dict, List, Optional

import torch
from torch import nn

from . import ops, torch_ops
from ._cabi import NanovsError

# ----------------------------------------------------------------------------------------------
# config letters (kp2dtiny.py:46-218)
# ----------------------------------------------------------------------------------------------
_S = [16, 32, 32, 64, 64, 128]
_N = [16, 24, 24, 48, 48, 96]
_D = [64, 128, 128, 256, 256, 512]

KP2DTINY_CONFIGS = {
    "S": dict(nfeatures=32, channel_dims=_S, downsample=2, use_attention=False, leaky_relu=True, encoder_dim=64),
    "S_A": dict(nfeatures=32, channel_dims=_S, downsample=2, use_attention=True, leaky_relu=True, encoder_dim=64),
    "N": dict(nfeatures=32, channel_dims=_N, downsample=2, use_attention=False, leaky_relu=True, num_clusters=32,
              encoder_dim=48),
    "N_A": dict(nfeatures=32, channel_dims=_N, downsample=2, use_attention=True, leaky_relu=True, num_clusters=32,
                encoder_dim=48),
    "D": dict(nfeatures=128, channel_dims=_D, downsample=2, use_attention=True, leaky_relu=True, encoder_dim=128,
              global_descriptor_method="convap"),
    "F": dict(nfeatures=64, channel_dims=[16, 32, 64, 128, 128, 256], downsample=3, use_attention=False,
              leaky_relu=True),
    "GEM_N": dict(nfeatures=32, channel_dims=_N, downsample=2, use_attention=False, leaky_relu=True,
                  num_clusters=32, encoder_dim=48, global_descriptor_method="gem"),
    "GEM_S_A": dict(nfeatures=32, channel_dims=_S, downsample=2, use_attention=True, leaky_relu=True,
                    encoder_dim=64, global_descriptor_method="gem"),
    "CONVAP_S_A": dict(nfeatures=32, channel_dims=_S, downsample=2, use_attention=True, leaky_relu=True,
                       encoder_dim=64, global_descriptor_method="convap"),
}

KP2DTINYV3_CONFIGS = {
    "S": dict(nfeatures=32, channel_dims=_S, bn_momentum=0.1, downsample=2, use_attention=False, leaky_relu=True,
              encoder_dim=64),
    "S_A": dict(nfeatures=32, channel_dims=_S, bn_momentum=0.1, downsample=2, use_attention=True, leaky_relu=True,
                encoder_dim=64),
    "N": dict(nfeatures=32, channel_dims=_N, bn_momentum=0.1, downsample=2, use_attention=False, encoder_dim=48),
    "N_A": dict(nfeatures=32, channel_dims=_N, bn_momentum=0.1, downsample=2, use_attention=True, encoder_dim=48),
    "D": dict(nfeatures=128, channel_dims=_D, downsample=2, use_attention=False, leaky_relu=True, encoder_dim=128,
              global_descriptor_method="convap"),
    "D_A": dict(nfeatures=128, channel_dims=_D, downsample=2, use_attention=True, leaky_relu=True,
                encoder_dim=128, global_descriptor_method="convap"),
    "CONVAP_S_A": dict(nfeatures=32, channel_dims=_S, bn_momentum=0.1, downsample=2, use_attention=True,
                       leaky_relu=True, encoder_dim=64, global_descriptor_method="convap"),
}


def get_config(config, to_mcu=False, to_export=False, v3=False):
    """kp2dtiny.py:245-281.  Returns a *copy* (the reference mutates the shared dict, :271-278)."""
    table = KP2DTINYV3_CONFIGS if v3 else KP2DTINY_CONFIGS
    if config not in table:
        raise ValueError("Config {} not supported, choose from ".format(config), list(table.keys()))
    conf = copy.deepcopy(table[config])
    if to_mcu:
        conf["upscale_method"] = "convtranspose"
        conf["leaky_relu"] = False
    if to_export:
        conf["remove_netvlad"] = True
    return conf


def tiny_factory(config, n_classes, to_mcu=False, to_export=False, v3=False):
    """kp2dtiny.py:221-242."""
    conf = get_config(config, to_mcu=to_mcu, to_export=to_export, v3=v3)
    return (KP2DTinyV3 if v3 else KP2DTinyV2)(**conf, nClasses=n_classes)


# ----------------------------------------------------------------------------------------------
# parameter holders (same attribute names as the reference modules => same state_dict keys)
# ----------------------------------------------------------------------------------------------
class _ConvBnAct(nn.Module):
    """modules/base.py:14-46 (conv.weight, bn.{weight,bias,running_mean,running_var,num_batches_tracked})."""

    def __init__(self, c0, c1, bn_momentum=0.1):
        super().__init__()
        self.conv = nn.Conv2d(c0, c1, 3, 1, 1, bias=False)
        self.bn = nn.BatchNorm2d(c1, momentum=bn_momentum)


class _BackBone(nn.Module):  # modules/encoders.py:5-103
    def __init__(self, c0, c1, c2, c3, c4, bn_momentum):
        super().__init__()
        self.conv1a = _ConvBnAct(c0, c1, bn_momentum)
        self.conv1b = _ConvBnAct(c1, c2, bn_momentum)
        self.conv2a = _ConvBnAct(c2, c2, bn_momentum)
        self.conv2b = _ConvBnAct(c2, c3, bn_momentum)
        self.conv3a = _ConvBnAct(c3, c3, bn_momentum)
        self.conv3b = _ConvBnAct(c3, c4, bn_momentum)
        self.conv4a = _ConvBnAct(c4, c4, bn_momentum)
        self.conv4b = _ConvBnAct(c4, c4, bn_momentum)


class _TaskHead(nn.Module):  # modules/decoders/heads.py:7-35
    def __init__(self, c_in, c_hidden, c_out, bn_momentum):
        super().__init__()
        self.convDa = _ConvBnAct(c_in, c_hidden, bn_momentum)
        self.convDb = nn.Conv2d(c_hidden, c_out, 3, 1, 1)


class _TransposedConvUp(nn.Module):  # modules/base.py:80-117 (to_mcu upsampling; parameter holder)
    def __init__(self, c, bn_momentum=0.1):
        super().__init__()
        self.transposed_conv = nn.ConvTranspose2d(c, c // 4, kernel_size=3, stride=2, padding=1, output_padding=1,
                                                  bias=False)
        self.bn = nn.BatchNorm2d(c // 4, momentum=bn_momentum)

    def equivalent_conv(self, eps: float = 1e-5):
        """(weight (c, c, 3, 3), bias (c,)) of the 3x3 / pad-1 conv whose PixelShuffle(2) equals this layer with the
        BatchNorm folded in.  Stride-2 / k=3 / p=1 / op=1 transposed conv: out[2y+i, 2x+j] sums in[y+dy, x+dx] *
        W[k(i,dy), k(j,dx)] with k(0,0) = 1, k(1,0) = 2, k(1,1) = 0 (even outputs have a single tap); the 2x2
        footprint sits at rows/cols 1..2 of a centred 3x3 kernel, output channel co*4 + 2i + j is sub-pixel (i, j)."""
        W = self.transposed_conv.weight.detach().to(torch.float64)  # (cin, c/4, 3, 3)
        cin, c4 = W.shape[:2]
        bn = self.bn
        sc = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + eps)
        sh = bn.bias.detach().double() - bn.running_mean.detach().double() * sc
        W3 = torch.zeros(c4 * 4, cin, 3, 3, dtype=torch.float64, device=W.device)
        kmap = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
        rows = torch.arange(c4, device=W.device) * 4
        for (i, dy), ky in kmap.items():
            for (j, dx), kx in kmap.items():
                W3[rows + 2 * i + j, :, 1 + dy, 1 + dx] = (W[:, :, ky, kx] * sc.view(1, -1)).t()
        return W3.to(torch.float32), sh.repeat_interleave(4).to(torch.float32)


class _UpscaleHead(nn.Module):  # modules/decoders/heads.py:38-104
    def __init__(self, c0, c1, c2, c3, c4, c5, bn_momentum, upscale_method="pixelshuffle"):
        super().__init__()
        if upscale_method == "convtranspose":  # defined first in the reference (heads.py:53-56): state_dict order
            self.upsample = _TransposedConvUp(c2, bn_momentum)
        self.convA = _ConvBnAct(c0, c1, bn_momentum)
        self.convB = nn.Conv2d(c1, c2, 3, 1, 1)
        self.confAa = _ConvBnAct(c3, c4, bn_momentum)
        self.confBb = nn.Conv2d(c4, c5, 3, 1, 1)


class _ChanLN(nn.Module):  # modules/segformer.py:63-73
    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))
        self.b = nn.Parameter(torch.zeros(1, dim, 1, 1))


class _PreNorm(nn.Module):  # modules/segformer.py:76-83
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = _ChanLN(dim)


class _ESA(nn.Module):  # modules/segformer.py:86-98
    def __init__(self, dim, heads=4, reduction_ratio=2):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Conv2d(dim, dim, 1, bias=False)
        self.to_kv = nn.Conv2d(dim, dim * 2, reduction_ratio, stride=reduction_ratio, bias=False)
        self.to_out = nn.Conv2d(dim, dim, 1, bias=False)


class _DsConv(nn.Module):  # modules/segformer.py:43-60
    def __init__(self, dim):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(dim, dim, 3, padding=1, groups=dim), nn.Conv2d(dim, dim, 1))


class _MixFFN(nn.Module):  # modules/segformer.py:182-200
    def __init__(self, dim, expansion_factor=2):
        super().__init__()
        hidden = dim * expansion_factor
        self.net = nn.Sequential(nn.Conv2d(dim, hidden, 1), _DsConv(hidden), nn.GELU(), nn.Conv2d(hidden, dim, 1))


class _AttModule(nn.Module):  # modules/segformer.py:209-220
    def __init__(self, c):
        super().__init__()
        self.att = _PreNorm(c, _ESA(c))
        self.mff = _PreNorm(c, _MixFFN(c))


class _SegHead(nn.Module):
    """The four segmentation heads (modules/decoders/segmentation.py:8,169,350,478) as one holder."""

    def __init__(self, c_in, c_hidden, c_exp, c_out, d1, bn_momentum, attention, n_feat=None, depth=False,
                 upscale_method="pixelshuffle"):
        super().__init__()
        fused = n_feat is not None  # V3: seg + feat from one trunk
        self.dim_split = c_hidden // 2
        self.depth = bool(depth) and fused
        # V3 with depth: the trunk's last block is three slices wide, the middle one feeds featD
        # (segmentation.py:187-191, 515-519)
        self.c_last = c_hidden + (self.dim_split if self.depth else 0)
        if fused:
            assert c_hidden % 2 == 0, "c_hidden must be divisible by 2"
        cb = lambda a, b: _ConvBnAct(a, b, bn_momentum)  # noqa: E731
        if attention:
            layers = [cb(c_in, c_hidden), _AttModule(c_hidden), _AttModule(c_hidden)]
        else:
            layers = [cb(c_in, c_hidden), cb(c_hidden, c_hidden), cb(c_hidden, c_hidden), cb(c_hidden, c_hidden)]
        layers += [cb(c_hidden, d1), cb(c_hidden + d1 // 4, c_hidden), cb(c_hidden, d1), cb(c_exp, self.c_last),
                   nn.Conv2d(self.dim_split if fused else c_hidden, c_out, 3, 1, 1)]
        self.convs = nn.ModuleList(layers)
        if fused:
            self.featB = nn.Conv2d(self.dim_split, n_feat, 3, 1, 1)
        if self.depth:
            self.featD = nn.Conv2d(self.dim_split, 1, 3, 1, 1, bias=False)  # segmentation.py:284-287
        if upscale_method == "convtranspose":  # segmentation.py:116-118, 295-297, 432-434, 569-571
            self.upsample = _TransposedConvUp(d1, bn_momentum)
            self.upsample2 = _TransposedConvUp(d1, bn_momentum)

    def freeze(self, except_last_layer=False):  # segmentation.py:159-166
        for p in self.parameters():
            p.requires_grad = False
        if except_last_layer:
            for p in self.convs[len(self.convs) - 1].parameters():
                p.requires_grad = True


class _NetVLAD(nn.Module):  # modules/aggregators/netvlad.py:19-48
    def __init__(self, num_clusters, dim):
        super().__init__()
        self.num_clusters, self.dim = num_clusters, dim
        self.conv = nn.Conv2d(dim, num_clusters, kernel_size=(1, 1), bias=False)
        self.centroids = nn.Parameter(torch.rand(num_clusters, dim))

    def get_desc_size(self):
        return self.dim * self.num_clusters

    def init_params(self, clsts, traindescs):  # netvlad.py:50-63 (vladv2=False branch)
        import numpy as np

        assign = clsts / np.linalg.norm(clsts, axis=1, keepdims=True)
        dots = np.dot(assign, traindescs.T)
        dots.sort(0)
        dots = dots[::-1, :]
        alpha = (-np.log(0.01) / np.mean(dots[0, :] - dots[1, :])).item()
        dev = self.centroids.device
        self.centroids = nn.Parameter(torch.from_numpy(clsts).to(dev))
        self.conv.weight = nn.Parameter(torch.from_numpy(alpha * assign).unsqueeze(2).unsqueeze(3).to(dev))


class _GeM(nn.Module):  # modules/aggregators/gem.py:8-33 (parameter holder; the math is nvs_gem)
    def __init__(self, c, p=3, eps=1e-6, unshuffle=4):
        super().__init__()
        if unshuffle != 4:
            raise NotImplementedError("GeM: VPRHead always uses unshuffle=4 (vpr.py:71)")
        self.p = nn.Parameter(torch.ones(1) * p)
        self.eps = eps
        self.f = unshuffle * unshuffle

    def get_factor(self):
        return self.f


class _ConvAP(nn.Module):  # modules/aggregators/convap.py:19-27 (parameter holder; the math is nvs_convap)
    def __init__(self, in_channels, out_channels=512, s1=2, s2=2):
        super().__init__()
        self.channel_pool = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=True)
        self.s1, self.s2 = s1, s2


class _VPRHead(nn.Module):  # modules/decoders/vpr.py:8-76
    def __init__(self, c_in, encoder_dim, num_clusters, bn_momentum, remove_netvlad, method):
        super().__init__()
        self.convlad1 = _ConvBnAct(c_in, encoder_dim, bn_momentum)
        self.convlad2 = _ConvBnAct(encoder_dim, encoder_dim, bn_momentum)
        self.convlad3 = _ConvBnAct(encoder_dim, encoder_dim, bn_momentum)
        self.remove_netvlad = remove_netvlad
        self.method = method
        if method == "netvlad":
            if not remove_netvlad:
                self.netvlad = _NetVLAD(num_clusters, encoder_dim)
                self.global_desc_dim = self.netvlad.get_desc_size()
            else:
                self.global_desc_dim = 0
        elif method == "gem":  # vpr.py:70-72
            self.netvlad = _GeM(encoder_dim, unshuffle=4)
            self.global_desc_dim = encoder_dim * self.netvlad.get_factor()
        elif method == "convap":  # vpr.py:73-76
            self.netvlad = _ConvAP(encoder_dim, encoder_dim, 4, 4)
            self.global_desc_dim = encoder_dim * 16
        else:
            raise ValueError(f"global_descriptor_method={method!r}")


def _p32(c: int) -> int:
    """Channel count of the zero-padded channels-last buffer that carries ``c`` real channels."""
    return (c + 31) // 32 * 32


# ----------------------------------------------------------------------------------------------
# launch plan
# ----------------------------------------------------------------------------------------------
class _Plan:
    """Everything shape dependent: cached intermediates + pre-filled NvsConvArgs, replayed per forward."""

    def __init__(self, model: "_KP2DTinyBase", B: int, H: int, W: int, device: torch.device):
        self.steps: List = []
        self.out_slots: Dict[str, List] = {}
        self.B, self.H, self.W, self.device = B, H, W, device
        self.tc_slice = getattr(model, "tc_slice", 128)  # output-channel slice width of wide tensor-core convs
        self.bufs: Dict[str, torch.Tensor] = {}
        self.meta: Dict[int, dict] = {}   # step index -> {flops, bytes} of conv launches (for bench.py)
        self.profile: Optional[dict] = None  # {"idx": step index, "events": [(start, end), ...]}
        self.graph = None                 # CUDA graph of the whole launch sequence (small batches)
        self.graph_x = None
        self.graph_outs: Optional[Dict[str, torch.Tensor]] = None
        self.runs = 0
        self.in_mode = ops.IN_PLAIN       # how the stem layer reads its input (IN_PLAIN / IN_U8_HWC / IN_UNIT)
        model._build_plan(self)

    def buf(self, name: str, c: int, h: int, w: int) -> torch.Tensor:
        t = torch.empty(self.B, c, h, w, device=self.device, dtype=torch.float32)
        self.bufs[name] = t
        return t

    def conv(self, packed, src0, cout, *, out_name=None, **kw):
        """Register a conv; if ``out_name`` is given its dst pointer is patched per call (fresh output)."""
        wp, bp = packed
        args = ops.make_conv_args(src0, wp, bp, cout, **kw)
        cin = (4 * args.c0) if args.in_mode == ops.IN_S2D else (args.c0 + args.c1)
        n_out = cout * args.H * args.W
        if args.out_mode == ops.OUT_POOL:
            n_out //= 4
        elif args.out_mode == ops.OUT_BOTH:
            n_out += n_out // 4
        self.meta[len(self.steps)] = {
            "flops": 2.0 * args.ksize * args.ksize * cin * cout * args.H * args.W * self.B,
            "bytes": 4.0 * self.B * ((args.c0 + args.c1) * args.in_H * args.in_W + n_out),
            "shape": f"{cin}->{cout} k{args.ksize} @{args.H}x{args.W}"}
        self.steps.append(("conv", args))
        if out_name is not None:
            self.out_slots.setdefault(out_name, []).append(args)

    def buf_nhwc(self, name: str, c: int, h: int, w: int) -> torch.Tensor:
        # zero-initialised: padding channels no launch writes (N letters) are multiplied by zero weights, which only
        # gives zero if they are finite -- and the fp16 operands of the 3xFP16 convs overflow above 65504
        t = torch.zeros(self.B, h, w, c, device=self.device, dtype=torch.float32)
        self.bufs[name] = t
        return t

    def tc(self, packed, src0, cout, *, out_name=None, out2_name=None, **kw):
        """Register a tensor-core conv (channels-last operands).  ``out_name`` / ``out2_name``: forward outputs
        that receive dst / the second output of this launch."""
        op = ops.tc_conv(src0, packed, cout, slice_width=self.tc_slice, **kw)
        n_in = src0.shape[1] * src0.shape[2] * ((kw.get("c0") or src0.shape[3]) + (kw.get("c1") or (
            kw["src1"].shape[3] if kw.get("src1") is not None else 0)))
        n_out = cout * src0.shape[1] * src0.shape[2] * (1 if kw.get("dst_mode", 1) else 0)
        if kw.get("dst_pool") is not None:
            n_out += cout * (src0.shape[1] // 2) * (src0.shape[2] // 2)
        self.meta[len(self.steps)] = {"flops": op.flops, "bytes": 4.0 * self.B * (n_in + n_out), "shape": op.shape}
        self.steps.append(("tc", op, out_name, out2_name))

    def call(self, fn, *a):
        self.steps.append(("call", fn, a))


class _KP2DTinyBase(nn.Module):
    version = 0

    # --- shared constructor tail ---------------------------------------------------------------
    def _finish_init(self):
        self.cell = pow(2, self.downsample)  # kp2dtiny.py:455
        self.cross_ratio = 2.0  # :339
        self.global_desc_dim = self.vlad_head.global_desc_dim
        self.training = True  # the reference leaves construction in "training" (:456); callers set False
        self._packed = None
        self._packed_key = None
        self._plans: Dict = {}
        self._handle = None  # torch.ops.nanovs.kp2dtiny_forward handle (torch_ops.register_model)
        if self.downsample not in (2, 3):
            raise NotImplementedError("downsample must be 2 (cell 4) or 3 (cell 8, letter F)")
        if self.upscale_method not in ("pixelshuffle", "convtranspose"):
            raise NotImplementedError("Upscale method not implemented")  # heads.py:58, segmentation.py:120
        if self.use_attention and self.channel_dims[4] // 4 not in (12, 16, 64):
            raise NotImplementedError("attention seg head: head_dim 12 / 16 / 64 (letters S_A, N_A, D, D_A)")
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())
        # conv backend: "tc" = tcgen05 3xTF32 implicit GEMM on channels-last maps (needs 32-channel multiples:
        # the S letters), "ffma" = exact fp32 direct conv (any channel count: the N letters).
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        # (channel counts that are not multiples of 32 -- the N letters: 24/48/72/96 -- run on zero-padded
        # 32-channel rows: padded weights are zero, so padded activations stay exactly zero)
        # 16-channel stem output: conv1b runs the paired-tap variant (32 output channels); otherwise the stem output
        # must fill whole 32-channel rows.  Layers wider than 128 channels run as several launches (ops.TcConvSplit).
        tc_ok = ((c1 == 16 and _p32(c2) <= 32) or c1 % 32 == 0) and d1 % 4 == 0 and (_p32(d1) <= 128 or d1 % 128 == 0)
        # The tensor core accumulates with round-toward-zero; the error grows with K = 9 * Cin.  The 256/512-channel
        # letters (D, D_A) measured 1.2e-4 .. 1.9e-4 with 128-wide launches and <= 8e-5 when every layer runs as
        # 64-wide slices (the 64-wide kernel keeps the small correction products in their own accumulator half), so
        # those letters use 64-wide slices: 4.4x the FFMA backend's frame rate, inside the 1e-4 tolerance.
        self.tc_slice = int(os.environ.get("NVS_TC_SLICE", "64" if max(c4, c5) > 128 else "128"))
        self.conv_backend = os.environ.get("NVS_CONV_BACKEND", "tc" if tc_ok else "ffma")
        # arithmetic of the tensor-core convs: "f16" (3xFP16, csrc/conv_rs.cu; needs |activation| < 65504: checked on the
        # first batch of every launch plan, with an automatic switch to "tf32") or "tf32" (3xTF32, csrc/conv_tc.cu)
        self.conv_math = ops.conv_math()
        # batches up to this size replay a captured CUDA graph (0 disables)
        self.cuda_graph_max_batch = int(os.environ.get("NVS_CUDA_GRAPH_MAX_BATCH", "16"))
        if self.conv_backend == "tc" and not tc_ok:
            raise NotImplementedError("tensor-core conv backend: unsupported channel configuration")

    # --- reference API ------------------------------------------------------------------------
    def gather_info(self):  # kp2dtiny.py:463-485
        params = inspect.signature(self.__init__).parameters
        init_args = {n: getattr(self, n) for n in params.keys() if hasattr(self, n)}
        total = sum(p.numel() for p in self.parameters())
        train = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {"init_args": init_args, "total_params": total, "trainable_params": train,
                "netvlad_dim": self.global_desc_dim, "upscale_method": self.upscale_method,
                "leaky_relu": self.leaky_relu, "use_attention": self.use_attention}

    def get_global_desc_dim(self):
        return self.global_desc_dim

    def get_netvlad_dim(self):
        return self.global_desc_dim

    def get_num_clusters(self):
        return self.vlad_head.netvlad.num_clusters

    def init_netvlad(self, clsts, traindescs):
        self.vlad_head.netvlad.init_params(clsts, traindescs)
        self._invalidate()

    def freeze_backbone(self):
        for p in self.backbone.parameters():
            p.requires_grad = False

    def freeze_segmentation(self, except_last_layer=False):
        self.seg_head.freeze(except_last_layer)

    def fuse(self):
        """Reference: torch.quantization.fuse_modules for PTQ (kp2dtiny.py:507-513).  BatchNorm is always
        folded at pack time here, so this is a no-op kept for API compatibility."""
        return None

    # --- cache handling -----------------------------------------------------------------------
    def _invalidate(self):
        self._packed = None
        self._plans = {}

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._invalidate()
        return r

    def repack(self):
        """Call after editing parameters in place (load_state_dict / .to() are tracked automatically)."""
        self._invalidate()

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _ensure_packed(self, device):
        key = self._param_key()
        if self._packed is None or self._packed_key != key:
            first = next(self.parameters())
            if first.device != device:
                raise NanovsError(f"model parameters are on {first.device} but the input is on {device}; "
                                  "call model.to(device) first")
            with torch.no_grad():
                self._packed = self._pack()
            self._packed_key = key
            self._plans = {}
        return self._packed

    def _pk_block(self, m: _ConvBnAct, tc: bool = False, seg=None):
        bn = {"weight": m.bn.weight, "bias": m.bn.bias, "running_mean": m.bn.running_mean,
              "running_var": m.bn.running_var}
        if tc:
            return ops.pack_conv_tc(m.conv.weight, bn=bn, eps=m.bn.eps, cin_segments=self._segs(m.conv, seg),
                                    math=self.conv_math)
        return ops.pack_conv(m.conv.weight, bn=bn, eps=m.bn.eps)

    def _pk_block_pair(self, ma: _ConvBnAct, mb: _ConvBnAct):
        """Two conv+BN blocks that read the SAME input as ONE tensor-core conv with their output channels side by
        side (each padded to a multiple of 32): the kernel's cost per pipeline step is nearly flat in N up to 128
        (DESIGN 5.1), so 64 -> 2x64 costs ~1.4x one 64 -> 64 conv instead of 2x."""
        ws, bs = [], []
        for m in (ma, mb):
            bn = {"weight": m.bn.weight, "bias": m.bn.bias, "running_mean": m.bn.running_mean,
                  "running_var": m.bn.running_var}
            w, b = ops._fold(m.conv.weight, None, bn, m.bn.eps)
            cp = _p32(w.shape[0])
            ws.append(torch.cat([w, w.new_zeros(cp - w.shape[0], *w.shape[1:])], 0))
            bs.append(torch.cat([b, b.new_zeros(cp - b.shape[0])], 0))
        return ops.pack_conv_tc(torch.cat(ws, 0), bias=torch.cat(bs, 0), cin_segments=self._segs(ma.conv, None),
                                math=self.conv_math)

    def _merge_head_convs(self) -> bool:
        """The first convs of the heads all read the backbone output: pair them up when two fit one 128-wide conv."""
        c4p = _p32(self.channel_dims[3])
        return self.conv_backend == "tc" and 2 * c4p <= 128 and _p32(self.encoder_dim) == c4p

    def _pk_conv(self, m: nn.Conv2d, s2d=False, tc: bool = False, seg=None):
        if tc:
            return ops.pack_conv_tc(m.weight, bias=m.bias, cin_segments=self._segs(m, seg), math=self.conv_math)
        return ops.pack_conv(m.weight, bias=m.bias, s2d=s2d)

    @staticmethod
    def _segs(conv: nn.Conv2d, seg):
        """Input-channel segments [(real, padded)] of a tensor-core conv; default: one segment padded to 32s."""
        cin = conv.weight.shape[1]
        if seg is None:
            return [(cin, cin if cin == 16 else _p32(cin))]
        assert sum(seg) == cin, (seg, cin)
        return [(c, _p32(c)) for c in seg]

    def _pack_att(self, m: _AttModule):
        f, g = m.att.fn, m.mff.fn.net
        return {
            "ln1": (m.att.norm.g.detach().reshape(-1).contiguous(), m.att.norm.b.detach().reshape(-1).contiguous()),
            "q": self._pk_conv(f.to_q), "kv": self._pk_conv(f.to_kv, s2d=True), "out": self._pk_conv(f.to_out),
            "ln2": (m.mff.norm.g.detach().reshape(-1).contiguous(), m.mff.norm.b.detach().reshape(-1).contiguous()),
            "m0": self._pk_conv(g[0]),
            "dw": (g[1].net[0].weight.detach().reshape(-1, 9).contiguous(), g[1].net[0].bias.detach().contiguous()),
            "pw": self._pk_conv(g[1].net[1]), "m3": self._pk_conv(g[3]), "heads": f.heads,
        }

    def _pk_up(self, m: "_TransposedConvUp", tc: bool):
        """to_mcu upsampling layer as its equivalent 3x3 conv (+ PixelShuffle epilogue), BatchNorm folded."""
        w3, b3 = m.equivalent_conv()
        if tc:
            return ops.pack_conv_tc(w3, bias=b3, cin_segments=[(w3.shape[1], _p32(w3.shape[1]))], math=self.conv_math)
        return ops.pack_conv(w3, bias=b3)

    def _pack_seg(self, P: dict, sh: "_SegHead", pre: str, tc: bool):
        if self.upscale_method == "convtranspose":
            P[pre + ".up1"] = self._pk_up(sh.upsample, tc)
            P[pre + ".up2"] = self._pk_up(sh.upsample2, tc)
        c1_, c2_, c3_, c4_, c5_, d1_ = self.channel_dims
        n_convs = len(sh.convs)
        for i, m in enumerate(sh.convs):
            # the two concat convs read [pixel-shuffled d1/4 | x (c4)] and [pixel-shuffled d1/4 | skip (c4)]
            seg = [d1_ // 4, c4_] if i in (n_convs - 4, n_convs - 2) else None
            if isinstance(m, _ConvBnAct):
                P[f"{pre}.{i}"] = self._pk_block(m, tc=tc, seg=seg)
            elif isinstance(m, _AttModule):
                P[f"{pre}.{i}"] = self._pack_att(m)
            else:
                P[f"{pre}.{i}"] = self._pk_conv(m, tc=tc)

    def _pack(self) -> dict:
        P = {}
        tc = self.conv_backend == "tc"
        # channels-last store mode of the kernels that feed tensor-core convs: 2 = split fp16 hi / lo format of the
        # 3xFP16 kernels (NVS_CONV_MATH=f16, the default), 1 = fp32 (3xTF32 kernels)
        self._nhwc_mode = 2 if self.conv_math == "f16" else 1
        bb = self.backbone
        for n in ("conv1a", "conv1b", "conv2a", "conv2b", "conv3a", "conv3b", "conv4a", "conv4b"):
            # the 3-channel stem layer stays on the FFMA kernel (K = 27 is too thin for a TMA row); conv1b
            # (16 channels) uses the 64-byte-row variant of the tensor-core kernel
            P["bb." + n] = self._pk_block(getattr(bb, n), tc=tc and n != "conv1a")
        self._pack_seg(P, self.seg_head, "seg", tc)
        vh = self.vlad_head
        for n in ("convlad1", "convlad2", "convlad3"):
            P["vlad." + n] = self._pk_block(getattr(vh, n), tc=tc)
        if not vh.remove_netvlad:
            nv = vh.netvlad
            if vh.method == "netvlad":
                P["vlad.assign"] = nv.conv.weight.detach().reshape(nv.num_clusters, nv.dim).contiguous().float()
                P["vlad.cent"] = nv.centroids.detach().contiguous().float()
            elif vh.method == "gem":
                P["vlad.gem_p"] = float(nv.p.detach().float().cpu())
            else:
                cp = nv.channel_pool
                P["vlad.cap_w"] = cp.weight.detach().reshape(cp.out_channels, cp.in_channels).contiguous().float()
                P["vlad.cap_b"] = cp.bias.detach().contiguous().float()
        self._pack_heads(P)
        return P

    # --- forward --------------------------------------------------------------------------------
    def _check_input(self, x):
        if self.training is not False:
            raise NanovsError("nano_vs_slam_b200 is inference only: call model.eval() and set model.training = False "
                              "as the reference callers do (eval_multitask.py:195-196, frontend.py:56-58)")
        u8 = isinstance(x, torch.Tensor) and x.dtype == torch.uint8
        if u8:
            # camera frames as they come off the decoder: uint8 (B,H,W,3).  /255 and (x-0.5)*2 (visual_odometry.py:283,
            # frontend.py:79) are applied by the stem kernel's load stage (tensor-core backend) or by
            # nvs_preprocess_u8 (FFMA backend); the reference model itself only ever sees the fp32 tensor.
            if x.dim() != 4 or x.shape[3] != 3:
                raise ValueError("uint8 input must be shaped (B,H,W,3)")
        elif not isinstance(x, torch.Tensor) or x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected a (B,3,H,W) tensor")
        if not x.is_cuda:
            raise NanovsError("input must be a CUDA tensor: the sm_100a kernels are the only implementation")
        B, H, W = (x.shape[0], x.shape[1], x.shape[2]) if u8 else (x.shape[0], x.shape[2], x.shape[3])
        hs, ws = (H // 2, W // 2) if self.downsample == 2 else (H // 2 // 2, W // 2 // 2)  # skip-level map
        if hs % 4 != 0 or ws % 4 != 0:
            # the reference fails in torch.cat when pool/pixel-shuffle sizes disagree
            raise RuntimeError(f"input {H}x{W}: the skip-level map {hs}x{ws} must have sides that are multiples "
                               "of 4 (pixel-shuffle/skip concat sizes would differ, as in the reference)")
        if u8:
            x = x.contiguous()
            return x if self._stem_fuses_input() else ops.preprocess_u8(x)
        return x.contiguous().float()

    def _forward_keys(self):
        """Order of the tensors returned by torch.ops.nanovs.kp2dtiny_forward."""
        return ("score", "coord", "feat", "vlad", "seg") + (("depth",) if self.depth else ())

    def _forward_shapes(self, B: int, H: int, W: int):
        """Output shapes of forward for a (B,3,H,W) input, in _forward_keys order (fake / meta implementation)."""
        H1, W1 = H // 2, W // 2
        H2, W2 = (H1, W1) if self.downsample == 2 else (H1 // 2, W1 // 2)
        H4, W4 = H2 // 2, W2 // 2
        vh = self.vlad_head
        vlad = (B, self.encoder_dim, H4, W4) if vh.remove_netvlad else (B, self.global_desc_dim)
        shapes = [(B, 1, H4, W4), (B, 2, H4, W4), (B, self.nfeatures, H2, W2), vlad, (B, self.nClasses, H2, W2)]
        if self.depth:
            shapes.append((B, 1, H2, W2))
        return shapes

    @torch.no_grad()
    def forward(self, x, unit_input: bool = False):
        """Returns {'score','coord','feat','vlad','seg'} like kp2dtiny.py:552-591 / :906-957.
        ``x``: (B,3,H,W) fp32 in [-1,1] as in the reference, or uint8 (B,H,W,3) camera frames.
        ``unit_input``: ``x`` is fp32 in [0,1] and the ``x.sub(0.5).mul(2.0)`` of the reference's own callers
        (frontend.py:79) is applied by the first kernel's load stage instead of a separate pass over the batch.
        The work is ONE custom operator, torch.ops.nanovs.kp2dtiny_forward (CUDA dispatch key only), whose
        implementation replays this module's launch plan through the C ABI."""
        if self.training is not False or not isinstance(x, torch.Tensor) or not x.is_cuda:
            self._check_input(x)  # raises the reference-facing error (training mode / CPU tensor / bad shape)
        if self._handle != id(self):  # first call, or a deepcopy that inherited the original's handle
            self._handle = torch_ops.register_model(self)
        outs = torch.ops.nanovs.kp2dtiny_forward(x, self._handle, bool(unit_input))
        return dict(zip(self._forward_keys(), outs))

    def _stem_fuses_input(self) -> bool:
        """The stem kernel (3 -> 16, tensor-core backend) converts uint8 / [0,1] frames in its load stage."""
        return self.conv_backend == "tc" and self.channel_dims[0] == 16

    @torch.no_grad()
    def _forward_impl(self, x, unit_input: bool = False):
        x = self._check_input(x)
        if x.dtype == torch.uint8:
            B, H, W, _ = x.shape
            in_mode = ops.IN_U8_HWC
        else:
            B, _, H, W = x.shape
            in_mode = ops.IN_PLAIN
            if unit_input and self._stem_fuses_input():
                in_mode = ops.IN_UNIT
            elif unit_input:
                x = x.sub(0.5).mul(2.0)  # the reference's own line (frontend.py:79); FFMA backend / wide stems only
        self._ensure_packed(x.device)
        # uint8 / [0,1] frames: own plan (own CUDA graph with the input mode baked in + own staging buffer)
        key = (B, H, W, x.device, in_mode)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:
                self._plans.clear()
            plan = self._plans[key] = _Plan(self, B, H, W, x.device)
            plan.in_mode = in_mode
        first = plan.runs == 0 and self.conv_backend == "tc" and self.conv_math == "f16"
        if first:
            ops.conv_rs_range_flag(reset=True)
        out = self._run(plan, x)
        if first and ops.conv_rs_range_flag(reset=True):  # (synchronises: once per launch plan)
            # an activation left fp16's range: the 3xFP16 operands of the following layer were not finite.  Switch
            # this model to the 3xTF32 kernels (fp32's exponent range) and run the batch again.
            import warnings
            warnings.warn("nano_vs_slam_b200: activations beyond the fp16 range (|x| >= 60000); this model now uses the "
                          "3xTF32 tensor-core convs (NVS_CONV_MATH=tf32)")
            self.conv_math = "tf32"
            self._invalidate()
            return self._forward_impl(x, unit_input=(in_mode == ops.IN_UNIT))
        return out

    def _launch_all(self, plan: _Plan, x: torch.Tensor, outs: Dict[str, torch.Tensor]) -> None:
        """Enqueue every kernel of the plan on the current stream (pure launches: nothing allocates or syncs)."""
        for name, slots in plan.out_slots.items():
            for a in slots:
                a.dst = outs[name].data_ptr()
        plan.in_args.src0 = x.data_ptr()
        plan.in_args.in_mode = plan.in_mode
        run_conv = ops.run_conv
        prof = plan.profile
        for i, st in enumerate(plan.steps):
            if prof is not None and i == prof["idx"]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if st[0] == "conv":
                    run_conv(st[1])
                elif st[0] == "tc":
                    st[1].run(outs[st[2]] if st[2] is not None else None, outs[st[3]] if st[3] is not None else None)
                else:
                    st[1](outs, *st[2])
                e1.record()
                prof["events"].append((e0, e1))
            elif st[0] == "conv":
                run_conv(st[1])
            elif st[0] == "tc":
                st[1].run(outs[st[2]] if st[2] is not None else None, outs[st[3]] if st[3] is not None else None)
            else:
                st[1](outs, *st[2])

    def _run(self, plan: _Plan, x: torch.Tensor):
        dev = x.device
        plan.runs += 1
        # Small batches are launch bound (~32 launches for <1 ms of GPU work): after two eager runs (which also
        # set the kernels' function attributes) the whole sequence is captured once into a CUDA graph with static
        # input/output buffers and replayed; results are copied out so callers still own fresh tensors.
        use_graph = (self.cuda_graph_max_batch > 0 and plan.B <= self.cuda_graph_max_batch and plan.profile is None
                     and plan.runs > 2)
        if use_graph:
            if plan.graph is None:
                plan.graph_x = torch.empty_like(x)
                plan.graph_outs = {n: torch.empty(sh, device=dev, dtype=torch.float32) for n, sh in plan.out_shapes.items()}
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                n0 = ops.LAUNCHES[0]
                with torch.cuda.graph(g):
                    self._launch_all(plan, plan.graph_x, plan.graph_outs)
                plan.graph_kernels = ops.LAUNCHES[0] - n0  # kernels recorded in the graph
                ops.LAUNCHES[0] = n0                       # capture launched nothing
                plan.graph = g
            plan.graph_x.copy_(x)
            plan.graph.replay()
            ops.LAUNCHES[0] += plan.graph_kernels
            outs = {n: t.clone() for n, t in plan.graph_outs.items()}
        else:
            outs = {name: torch.empty(shape, device=dev, dtype=torch.float32) for name, shape in plan.out_shapes.items()}
            self._launch_all(plan, x, outs)
        result = {"score": outs["score"], "coord": outs["coord"], "feat": outs["feat"]}
        if "vlad" in outs:
            result["vlad"] = outs["vlad"]
        else:
            result["vlad"] = plan.bufs["v3"].clone()  # remove_netvlad: raw encoder map (vpr.py:83-89)
        result["seg"] = outs["seg"]
        if "depth" in outs:
            result["depth"] = outs["depth"]  # already sigmoid-ed (kp2dtiny.py:589, :956)
        return result

    def _plan_aggregator(self, pl: _Plan, P, v3: torch.Tensor, B: int, enc: int, H4: int, W4: int):
        """Global-descriptor aggregator on the (B, enc, H/4, W/4) NCHW encoder map (vpr.py:84-88)."""
        vh = self.vlad_head
        if vh.remove_netvlad:
            return
        if vh.method == "netvlad":
            K = vh.netvlad.num_clusters
            pl.out_shapes["vlad"] = (B, K * enc)
            ws = torch.empty(ops.netvlad_workspace_bytes(B, enc, K, H4 * W4), dtype=torch.uint8, device=pl.device)
            pl.bufs["vlad_ws"] = ws
            pl.call(lambda outs, v3=v3, ws=ws: ops.netvlad(v3, P["vlad.assign"], P["vlad.cent"], out=outs["vlad"],
                                                            workspace=ws))
        elif vh.method == "gem":
            if H4 % 4 or W4 % 4:
                raise ValueError(f"GeM: PixelUnshuffle(4) needs the {H4}x{W4} encoder map to be a multiple of 4 "
                                 "(the reference raises here too, gem.py:23)")
            pl.out_shapes["vlad"] = (B, 16 * enc)
            pl.call(lambda outs, v3=v3: ops.gem(v3, P["vlad.gem_p"], vh.netvlad.eps, out=outs["vlad"]))
        else:
            pl.out_shapes["vlad"] = (B, 16 * enc)
            ws = torch.empty(int(ops.lib().nvs_convap_workspace_bytes(B, enc, 4, 4)), dtype=torch.uint8, device=pl.device)
            pl.bufs["vlad_ws"] = ws
            pl.call(lambda outs, v3=v3, ws=ws: ops.convap(v3, P["vlad.cap_w"], P["vlad.cap_b"], 4, 4, out=outs["vlad"],
                                                           workspace=ws))

    # --- plan construction ------------------------------------------------------------------------
    def _build_plan(self, pl: _Plan):
        if self.conv_backend == "tc":
            return self._build_plan_tc(pl)
        return self._build_plan_ffma(pl)

    def _build_plan_ffma(self, pl: _Plan):
        P = self._packed
        B, H, W = pl.B, pl.H, pl.W
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        act = ops.ACT_LRELU if self.leaky_relu else ops.ACT_RELU
        # downsample = 3 (letter F, cell 8) pools a second time after conv2b (encoders.py:116-117): from conv3a
        # on every map is one level smaller, the graph is the same.  (H2,W2) = skip level, (H4,W4) = cell level.
        H1, W1 = H // 2, W // 2
        H2, W2 = (H1, W1) if self.downsample == 2 else (H1 // 2, W1 // 2)
        H4, W4 = H2 // 2, W2 // 2
        H8, W8 = H4 // 2, W4 // 2
        nf, ncls = self.nfeatures, self.nClasses
        pl.out_shapes = {"score": (B, 1, H4, W4), "coord": (B, 2, H4, W4), "feat": (B, nf, H2, W2),
                         "seg": (B, ncls, H2, W2)}

        # ---- backbone (modules/encoders.py:105-129) ----
        xin = torch.empty(B, 3, H, W, device=pl.device)  # shape carrier; pointer patched per call
        t1a = pl.buf("t1a", c1, H, W)
        pl.conv(P["bb.conv1a"], xin, c1, act=act, dst=t1a)
        pl.in_args = pl.steps[-1][1]
        p1 = pl.buf("p1", c2, H1, W1)
        pl.conv(P["bb.conv1b"], t1a, c2, act=act, out_mode=ops.OUT_POOL, dst2=p1)
        t2a = pl.buf("t2a", c2, H1, W1)
        pl.conv(P["bb.conv2a"], p1, c2, act=act, dst=t2a)
        t2b = pl.buf("t2b", c3, H2, W2)
        if self.downsample == 3:
            pl.conv(P["bb.conv2b"], t2a, c3, act=act, out_mode=ops.OUT_POOL, dst2=t2b)
        else:
            pl.conv(P["bb.conv2b"], t2a, c3, act=act, dst=t2b)
        t3a = pl.buf("t3a", c3, H2, W2)
        pl.conv(P["bb.conv3a"], t2b, c3, act=act, dst=t3a)
        skip = pl.buf("skip", c4, H2, W2)
        p3 = pl.buf("p3", c4, H4, W4)
        pl.conv(P["bb.conv3b"], t3a, c4, act=act, out_mode=ops.OUT_BOTH, dst=skip, dst2=p3)
        t4a = pl.buf("t4a", c4, H4, W4)
        pl.conv(P["bb.conv4a"], p3, c4, act=act, dst=t4a)
        xb = pl.buf("xb", c4, H4, W4)
        pl.conv(P["bb.conv4b"], t4a, c4, act=act, dst=xb)

        # ---- keypoint / descriptor heads (version specific) ----
        self._plan_heads(pl, xb, skip, act)

        # ---- segmentation trunk (modules/decoders/segmentation.py) ----
        s7, last = self._plan_trunk_ffma(pl, "seg", "", self.seg_head, xb, skip, act)
        self._plan_seg_out(pl, s7, last)
        if self.depth and self.version == 2:  # second trunk, one output channel, sigmoid (kp2dtiny.py:588-590)
            d7, dlast = self._plan_trunk_ffma(pl, "dep", "_d", self.depth_head, xb, skip, act)
            pl.out_shapes["depth"] = (B, 1, H2, W2)
            pl.conv(dlast, d7, 1, act=ops.ACT_SIGMOID, dst=torch.empty(B, 1, H2, W2, device=pl.device),
                    out_name="depth")

        # ---- VPR head (modules/decoders/vpr.py:78-89) ----
        enc = self.encoder_dim
        v1 = pl.buf("v1", enc, H4, W4)
        pl.conv(P["vlad.convlad1"], xb, enc, act=act, dst=v1)
        v2 = pl.buf("v2", enc, H4, W4)
        pl.conv(P["vlad.convlad2"], v1, enc, act=act, dst=v2)
        v3 = pl.buf("v3", enc, H4, W4)
        pl.conv(P["vlad.convlad3"], v2, enc, act=act, dst=v3)
        self._plan_aggregator(pl, P, v3, B, enc, H4, W4)

    def _plan_trunk_ffma(self, pl: _Plan, pre: str, tag: str, head: "_SegHead", xb, skip, act):
        """Segmentation trunk up to its last conv block (segmentation.py:126-152 / 314-334 / 442-463 / 588-608);
        returns (last block's output, packed final conv).  ``pre`` selects the packed weights ("seg" / "dep"),
        ``tag`` keeps the buffers of a second trunk apart."""
        P = self._packed
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        _, _, H4, W4 = xb.shape
        H2, W2 = skip.shape[2:]
        H8, W8 = H4 // 2, W4 // 2
        s0 = pl.buf("s0" + tag, c5, H4, W4)
        pl.conv(P[pre + ".0"], xb, c5, act=act, dst=s0)
        sp3 = pl.buf("sp3" + tag, c5, H8, W8)
        if self.use_attention:
            sp = pl.buf("sp" + tag, c5, H8, W8)
            self._plan_att(pl, P[pre + ".1"], s0, c5, H4, W4, "a1" + tag, pooled_out=sp)
            self._plan_att(pl, P[pre + ".2"], sp, c5, H8, W8, "a2" + tag, plain_out=sp3)
            nxt = 3
        else:
            sp = pl.buf("sp" + tag, c5, H8, W8)
            pl.conv(P[pre + ".1"], s0, c5, act=act, out_mode=ops.OUT_POOL, dst2=sp)
            s2 = pl.buf("s2" + tag, c5, H8, W8)
            pl.conv(P[pre + ".2"], sp, c5, act=act, dst=s2)
            pl.conv(P[pre + ".3"], s2, c5, act=act, dst=sp3)
            nxt = 4
        ps1 = pl.buf("ps1" + tag, d1 // 4, H4, W4)
        mcu = self.upscale_method == "convtranspose"  # to_mcu: conv -> [transposed conv + BN + act] instead of a shuffle
        if mcu:
            u1 = pl.buf("u1" + tag, d1, H8, W8)
            pl.conv(P[f"{pre}.{nxt}"], sp3, d1, act=act, dst=u1)
            pl.conv(P[pre + ".up1"], u1, d1, act=act, out_mode=ops.OUT_SHUFFLE, dst=ps1)
        else:
            pl.conv(P[f"{pre}.{nxt}"], sp3, d1, act=act, out_mode=ops.OUT_SHUFFLE, dst=ps1)
        s5 = pl.buf("s5" + tag, c5, H4, W4)
        pl.conv(P[f"{pre}.{nxt + 1}"], ps1, c5, act=act, src1=xb, dst=s5)
        ps2 = pl.buf("ps2" + tag, d1 // 4, H2, W2)
        if mcu:
            u2 = pl.buf("u2" + tag, d1, H4, W4)
            pl.conv(P[f"{pre}.{nxt + 2}"], s5, d1, act=act, dst=u2)
            pl.conv(P[pre + ".up2"], u2, d1, act=act, out_mode=ops.OUT_SHUFFLE, dst=ps2)
        else:
            pl.conv(P[f"{pre}.{nxt + 2}"], s5, d1, act=act, out_mode=ops.OUT_SHUFFLE, dst=ps2)
        s7 = pl.buf("s7" + tag, head.c_last, H2, W2)
        pl.conv(P[f"{pre}.{nxt + 3}"], ps2, head.c_last, act=act, src1=skip, dst=s7)
        return s7, P[f"{pre}.{nxt + 4}"]

    def _build_plan_tc(self, pl: _Plan):
        """Same graph as _build_plan_ffma with channels-last intermediates and tcgen05 convs (csrc/conv_tc.cu).
        Only the 3->16 stem layer and the attention internals use other kernels.  Intermediate buffers carry
        channel counts padded to multiples of 32 (no-op for the S letters); padded channels are exact zeros."""
        P = self._packed
        B, H, W = pl.B, pl.H, pl.W
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        c2p, c3p, c4p, c5p, qp = _p32(c2), _p32(c3), _p32(c4), _p32(c5), _p32(d1 // 4)
        d1p = _p32(d1)
        act = ops.ACT_LRELU if self.leaky_relu else ops.ACT_RELU
        H1, W1 = H // 2, W // 2
        H2, W2 = (H1, W1) if self.downsample == 2 else (H1 // 2, W1 // 2)  # skip level (see _build_plan_ffma)
        H4, W4 = H2 // 2, W2 // 2
        H8, W8 = H4 // 2, W4 // 2
        nf, ncls = self.nfeatures, self.nClasses
        pl.out_shapes = {"score": (B, 1, H4, W4), "coord": (B, 2, H4, W4), "feat": (B, nf, H2, W2),
                         "seg": (B, ncls, H2, W2)}
        # ---- stem layer on the FFMA kernel: NCHW in, channels-last out ----
        xin = torch.empty(B, 3, H, W, device=pl.device)
        t1a = pl.buf_nhwc("t1a", c1, H, W)
        pl.conv(P["bb.conv1a"], xin, c1, act=act, dst=t1a, dst_nhwc=self._nhwc_mode)
        pl.in_args = pl.steps[-1][1]
        # ---- backbone on tensor cores ----
        p1 = pl.buf_nhwc("p1", c2p, H1, W1)
        pl.tc(P["bb.conv1b"], t1a, c2p, act=act, dst=None, dst_mode=0, dst_pool=p1)
        t2a = pl.buf_nhwc("t2a", c2p, H1, W1)
        pl.tc(P["bb.conv2a"], p1, c2p, act=act, dst=t2a)
        t2b = pl.buf_nhwc("t2b", c3p, H2, W2)
        if self.downsample == 3:  # second pool after conv2b (encoders.py:116-117)
            pl.tc(P["bb.conv2b"], t2a, c3p, act=act, dst=None, dst_mode=0, dst_pool=t2b)
        else:
            pl.tc(P["bb.conv2b"], t2a, c3p, act=act, dst=t2b)
        t3a = pl.buf_nhwc("t3a", c3p, H2, W2)
        pl.tc(P["bb.conv3a"], t2b, c3p, act=act, dst=t3a)
        skip = pl.buf_nhwc("skip", c4p, H2, W2)
        p3 = pl.buf_nhwc("p3", c4p, H4, W4)
        pl.tc(P["bb.conv3b"], t3a, c4p, act=act, dst=skip, dst_pool=p3)
        t4a = pl.buf_nhwc("t4a", c4p, H4, W4)
        pl.tc(P["bb.conv4a"], p3, c4p, act=act, dst=t4a)
        xb = pl.buf_nhwc("xb", c4p, H4, W4)
        pl.tc(P["bb.conv4b"], t4a, c4p, act=act, dst=xb)

        self._plan_heads_tc(pl, xb, skip, act)

        # ---- segmentation trunk ----
        s7, last = self._plan_trunk_tc(pl, "seg", "", self.seg_head, xb, skip, act)
        self._plan_seg_out_tc(pl, s7, last)
        if self.depth and self.version == 2:  # second trunk, one output channel, sigmoid (kp2dtiny.py:588-590)
            d7, dlast = self._plan_trunk_tc(pl, "dep", "_d", self.depth_head, xb, skip, act)
            pl.out_shapes["depth"] = (B, 1, H2, W2)
            pl.tc(dlast, d7, 1, act=ops.ACT_SIGMOID, dst=None, dst_layout=1, dst_c_total=1, out_name="depth")

        # ---- VPR head ----
        enc = self.encoder_dim
        encp = _p32(enc)
        v2 = pl.buf_nhwc("v2", encp, H4, W4)
        if "v1_in_dva" in pl.bufs:  # convlad1 ran merged with a head conv (see _plan_heads_tc): channels [encp, 2*encp)
            pl.tc(P["vlad.convlad2"], pl.bufs["v1_in_dva"], encp, c0_off=encp, c0=encp, act=act, dst=v2)
        else:
            v1 = pl.buf_nhwc("v1", encp, H4, W4)
            pl.tc(P["vlad.convlad1"], xb, encp, act=act, dst=v1)
            pl.tc(P["vlad.convlad2"], v1, encp, act=act, dst=v2)
        v3 = pl.buf("v3", enc, H4, W4)  # NCHW, real channels, for the NetVLAD kernel
        pl.tc(P["vlad.convlad3"], v2, enc, act=act, dst=v3, dst_layout=1)
        self._plan_aggregator(pl, P, v3, B, enc, H4, W4)

    def _plan_trunk_tc(self, pl: _Plan, pre: str, tag: str, head: "_SegHead", xb, skip, act):
        """_plan_trunk_ffma on channels-last maps with tcgen05 convs."""
        P = self._packed
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        c5p, d1p, qp = _p32(c5), _p32(d1), _p32(d1 // 4)
        _, H4, W4, _ = xb.shape
        H2, W2 = skip.shape[1:3]
        H8, W8 = H4 // 2, W4 // 2
        sp3 = pl.buf_nhwc("sp3" + tag, c5p, H8, W8)
        if self.use_attention:
            s0 = pl.buf("s0" + tag, c5, H4, W4)  # NCHW, real channels: the attention block's LN / projections read planes
            pl.tc(P[pre + ".0"], xb, c5, act=act, dst=s0, dst_layout=1)
            sp = pl.buf("sp" + tag, c5, H8, W8)
            self._plan_att(pl, P[pre + ".1"], s0, c5, H4, W4, "a1" + tag, pooled_out=sp)
            sp3.zero_()  # the FFMA 1x1 writes the real channels only; padding must read as zero
            self._plan_att(pl, P[pre + ".2"], sp, c5, H8, W8, "a2" + tag, plain_out=sp3, plain_nhwc=True)
            nxt = 3
        else:
            s0 = pl.buf_nhwc("s0" + tag, c5p, H4, W4)
            pl.tc(P[pre + ".0"], xb, c5p, act=act, dst=s0)
            sp = pl.buf_nhwc("sp" + tag, c5p, H8, W8)
            pl.tc(P[pre + ".1"], s0, c5p, act=act, dst=None, dst_mode=0, dst_pool=sp)
            s2 = pl.buf_nhwc("s2" + tag, c5p, H8, W8)
            pl.tc(P[pre + ".2"], sp, c5p, act=act, dst=s2)
            pl.tc(P[pre + ".3"], s2, c5p, act=act, dst=sp3)
            nxt = 4
        ps1 = pl.buf_nhwc("ps1" + tag, qp, H4, W4)
        mcu = self.upscale_method == "convtranspose"
        if mcu:
            u1 = pl.buf_nhwc("u1" + tag, d1p, H8, W8)
            pl.tc(P[f"{pre}.{nxt}"], sp3, d1p, act=act, dst=u1)
            pl.tc(P[pre + ".up1"], u1, d1p, act=act, dst=ps1, dst_mode=2)
        else:
            pl.tc(P[f"{pre}.{nxt}"], sp3, d1p, act=act, dst=ps1, dst_mode=2)
        s5 = pl.buf_nhwc("s5" + tag, c5p, H4, W4)
        pl.tc(P[f"{pre}.{nxt + 1}"], ps1, c5p, act=act, src1=xb, dst=s5)
        ps2 = pl.buf_nhwc("ps2" + tag, qp, H2, W2)
        if mcu:
            u2 = pl.buf_nhwc("u2" + tag, d1p, H4, W4)
            pl.tc(P[f"{pre}.{nxt + 2}"], s5, d1p, act=act, dst=u2)
            pl.tc(P[pre + ".up2"], u2, d1p, act=act, dst=ps2, dst_mode=2)
        else:
            pl.tc(P[f"{pre}.{nxt + 2}"], s5, d1p, act=act, dst=ps2, dst_mode=2)
        clp = _p32(head.c_last)
        s7 = pl.buf_nhwc("s7" + tag, clp, H2, W2)
        pl.tc(P[f"{pre}.{nxt + 3}"], ps2, clp, act=act, src1=skip, dst=s7)
        return s7, P[f"{pre}.{nxt + 4}"]

    def _plan_att(self, pl: _Plan, A: dict, x: torch.Tensor, C: int, h: int, w: int, tag: str,
                  pooled_out: Optional[torch.Tensor] = None, plain_out: Optional[torch.Tensor] = None,
                  plain_nhwc: bool = False):
        """SegFormerAttentionModule (modules/segformer.py:209-220): LN -> q / kv -> attention -> to_out ->
        LN -> 1x1 -> dw3x3 -> 1x1 -> GELU -> 1x1.  No residual adds."""
        ln1 = pl.buf(tag + ".ln1", C, h, w)
        pl.call(lambda outs: ops.channel_layernorm(x, A["ln1"][0], A["ln1"][1], 1e-5, out=ln1))
        q = pl.buf(tag + ".q", C, h, w)
        pl.conv(A["q"], ln1, C, ksize=1, dst=q)
        kv = pl.buf(tag + ".kv", 2 * C, h // 2, w // 2)
        pl.conv(A["kv"], ln1, 2 * C, ksize=1, in_mode=ops.IN_S2D, dst=kv)
        at = pl.buf(tag + ".att", C, h, w)
        pl.call(lambda outs: ops.attention(q, kv, A["heads"], out=at))
        ao = pl.buf(tag + ".ao", C, h, w)
        pl.conv(A["out"], at, C, ksize=1, dst=ao)
        ln2 = pl.buf(tag + ".ln2", C, h, w)
        pl.call(lambda outs: ops.channel_layernorm(ao, A["ln2"][0], A["ln2"][1], 1e-5, out=ln2))
        m0 = pl.buf(tag + ".m0", 2 * C, h, w)
        pl.conv(A["m0"], ln2, 2 * C, ksize=1, dst=m0)
        m1 = pl.buf(tag + ".m1", 2 * C, h, w)
        pl.call(lambda outs: ops.dwconv3x3(m0, A["dw"][0], A["dw"][1], out=m1))
        m2 = pl.buf(tag + ".m2", 2 * C, h, w)
        pl.conv(A["pw"], m1, 2 * C, ksize=1, act=ops.ACT_GELU, dst=m2)
        if pooled_out is not None:
            pl.conv(A["m3"], m2, C, ksize=1, out_mode=ops.OUT_POOL, dst2=pooled_out)
        else:
            pl.conv(A["m3"], m2, C, ksize=1, dst=plain_out, dst_nhwc=self._nhwc_mode if plain_nhwc else 0)

    # --- post_processing (kp2dtiny.py:593-647 / :959-1015) ------------------------------------------
    @torch.no_grad()
    def post_processing(self, out, H, W):
        score, shift, feat = out["score"], out["coord"], out["feat"]
        if not score.is_cuda:
            raise NanovsError("post_processing needs CUDA tensors (no CPU fallback)")
        sample = self.training is False
        o_s, o_c, o_f = torch.ops.nanovs.decode(score, shift, feat if sample else None, H, W, self.cell,
                                                self.cross_ratio)
        if sample:
            seg = out["seg"]
            # V2: argmax(softmax(logits)) == argmax(logits); V3: forward already returned probabilities
            out["seg"] = torch.ops.nanovs.seg_argmax(seg, o_c if self.sample_segmentation else None, H, W)
            out["feat"] = o_f
        else:
            out["feat"] = feat
        out["coord"] = o_c
        out["score"] = o_s
        return out

    def only_encoder(self, x):
        """L2-normalised VPR encoder map (kp2dtiny.py:515-518, vpr.py:85-86)."""
        x = self._check_input(x)
        if x.dtype == torch.uint8:
            x = ops.preprocess_u8(x)
        self.forward(x)
        plan = self._plans[(x.shape[0], x.shape[2], x.shape[3], x.device, ops.IN_PLAIN)]
        return ops.l2norm_channels(plan.bufs["v3"])


class KP2DTinyV2(_KP2DTinyBase):
    """Dedicated-decoder model (kp2dtiny.py:284-647)."""

    version = 2

    def __init__(self, nfeatures=256, device="cpu", channel_dims=[32, 64, 128, 256, 256, 512], bn_momentum=0.1,
                 nClasses=8, num_clusters=64, downsample=3, use_attention=False, mem_efficient=False,
                 upscale_method="pixelshuffle", remove_netvlad=False, leaky_relu=True, depth=False,
                 encoder_dim=None, global_descriptor_method="netvlad", **kwargs):
        super().__init__()
        self.device = device
        self.with_drop = True
        self.nfeatures, self.downsample, self.nClasses = nfeatures, downsample, nClasses
        self.sample_segmentation = False
        self.use_attention, self.leaky_relu = use_attention, leaky_relu
        self.remove_netvlad, self.upscale_method, self.depth = remove_netvlad, upscale_method, depth
        self.num_clusters, self.global_descriptor_method = num_clusters, global_descriptor_method
        self.mem_efficient = mem_efficient  # same math, same keys (netvlad.py:110-197): one kernel serves both
        self.bn_momentum = bn_momentum
        self.channel_dims = list(channel_dims)
        c1, c2, c3, c4, c5, d1 = channel_dims
        self.encoder_dim = encoder_dim if encoder_dim is not None else c4
        self.backbone = _BackBone(3, c1, c2, c3, c4, bn_momentum)
        self.score_head = _TaskHead(c4, c4, 1, bn_momentum)
        self.loc_head = _TaskHead(c4, c4, 2, bn_momentum)
        self.desc_head = _UpscaleHead(c4, c4, c3 * 4, c3 + c4, c4, nfeatures, bn_momentum, upscale_method)
        self.seg_head = _SegHead(c4, c5, c4 + c3, nClasses, d1, bn_momentum, use_attention,
                                 upscale_method=upscale_method)
        if depth:  # a second segmentation head with one output channel (kp2dtiny.py:402-437)
            self.depth_head = _SegHead(c4, c5, c4 + c3, 1, d1, bn_momentum, use_attention,
                                       upscale_method=upscale_method)
        self.vlad_head = _VPRHead(c4, self.encoder_dim, num_clusters, bn_momentum, remove_netvlad,
                                  global_descriptor_method)
        self._finish_init()

    def _pack_heads(self, P):
        tc = self.conv_backend == "tc"
        if self.depth:
            self._pack_seg(P, self.depth_head, "dep", tc)
        if self._merge_head_convs():
            P["kp.a"] = self._pk_block_pair(self.score_head.convDa, self.loc_head.convDa)
            P["dv.a"] = self._pk_block_pair(self.desc_head.convA, self.vlad_head.convlad1)
        else:
            P["score.a"] = self._pk_block(self.score_head.convDa, tc=tc)
            P["loc.a"] = self._pk_block(self.loc_head.convDa, tc=tc)
        if tc:
            P["kp.b"] = ops.pack_head_pair_tc(self.score_head.convDb.weight, self.score_head.convDb.bias,
                                              self.loc_head.convDb.weight, self.loc_head.convDb.bias,
                                              cpad=_p32(self.channel_dims[3]), math=self.conv_math)
        else:
            P["score.b"] = self._pk_conv(self.score_head.convDb)
            P["loc.b"] = self._pk_conv(self.loc_head.convDb)
        d = self.desc_head
        if not self._merge_head_convs():
            P["desc.A"] = self._pk_block(d.convA, tc=tc)
        P["desc.B"] = self._pk_conv(d.convB, tc=tc)
        P["desc.Aa"] = self._pk_block(d.confAa, tc=tc, seg=[self.channel_dims[2], self.channel_dims[3]])
        P["desc.Bb"] = self._pk_conv(d.confBb, tc=tc)
        if self.upscale_method == "convtranspose":
            P["desc.up"] = self._pk_up(d.upsample, tc)

    def _plan_heads(self, pl, xb, skip, act):
        P = self._packed
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        B, _, H4, W4 = xb.shape
        H2, W2 = skip.shape[2:]
        sh = pl.buf("sh", c4, H4, W4)
        pl.conv(P["score.a"], xb, c4, act=act, dst=sh)
        pl.conv(P["score.b"], sh, 1, act=ops.ACT_SIGMOID, dst=torch.empty(B, 1, H4, W4, device=pl.device),
                out_name="score")
        lh = pl.buf("lh", c4, H4, W4)
        pl.conv(P["loc.a"], xb, c4, act=act, dst=lh)
        pl.conv(P["loc.b"], lh, 2, act=ops.ACT_TANH, dst=torch.empty(B, 2, H4, W4, device=pl.device),
                out_name="coord")
        # UpscaleHead (heads.py:91-104): convB's epilogue writes the pixel-shuffled map; confAa reads
        # [shuffled | skip] as two sources, so torch.cat is never materialised.
        da = pl.buf("da", c4, H4, W4)
        pl.conv(P["desc.A"], xb, c4, act=act, dst=da)
        dps = pl.buf("dps", c3, H2, W2)
        if self.upscale_method == "convtranspose":
            db = pl.buf("db", 4 * c3, H4, W4)
            pl.conv(P["desc.B"], da, 4 * c3, dst=db)
            pl.conv(P["desc.up"], db, 4 * c3, act=act, out_mode=ops.OUT_SHUFFLE, dst=dps)
        else:
            pl.conv(P["desc.B"], da, 4 * c3, out_mode=ops.OUT_SHUFFLE, dst=dps)
        dA = pl.buf("dA", c4, H2, W2)
        pl.conv(P["desc.Aa"], dps, c4, act=act, src1=skip, dst=dA)
        pl.conv(P["desc.Bb"], dA, self.nfeatures, dst=torch.empty(B, self.nfeatures, H2, W2, device=pl.device),
                out_name="feat")

    def _plan_seg_out(self, pl, s7, packed_last):
        B, _, H2, W2 = s7.shape
        pl.conv(packed_last, s7, self.nClasses, dst=torch.empty(B, self.nClasses, H2, W2, device=pl.device),
                out_name="seg")

    def _plan_heads_tc(self, pl, xb, skip, act):
        P = self._packed
        c1, c2, c3, c4, c5, d1 = self.channel_dims
        c3p, c4p = _p32(c3), _p32(c4)
        B, H4, W4, _ = xb.shape
        H2, W2 = skip.shape[1:3]
        merged = self._merge_head_convs()
        if merged:
            # score.convDa | loc.convDa and desc.convA | vlad.convlad1 as two 2*c4-wide convs of the backbone output
            kpa = pl.buf_nhwc("kpa", 2 * c4p, H4, W4)
            pl.tc(P["kp.a"], xb, 2 * c4p, act=act, dst=kpa)
            sh, lh, sl_off = kpa, kpa, c4p
            dva = pl.buf_nhwc("dva", 2 * c4p, H4, W4)
            pl.tc(P["dv.a"], xb, 2 * c4p, act=act, dst=dva)
            da = dva
            pl.bufs["v1_in_dva"] = dva  # the VPR head continues from channels [c4p, 2*c4p) of this buffer
        else:
            sh = pl.buf_nhwc("sh", c4p, H4, W4)
            pl.tc(P["score.a"], xb, c4p, act=act, dst=sh)
            lh = pl.buf_nhwc("lh", c4p, H4, W4)
            pl.tc(P["loc.a"], xb, c4p, act=act, dst=lh)
            sl_off = 0
            da = pl.buf_nhwc("da", c4p, H4, W4)
            pl.tc(P["desc.A"], xb, c4p, act=act, dst=da)
        # both 1- and 2-channel output convs as one tensor-core launch (block-diagonal weight over [sh | lh]),
        # sigmoid / tanh and the split into the two NCHW outputs happen in its epilogue
        pl.tc(P["kp.b"], sh, 3, src1=lh, c0_off=0, c0=c4p, c1_off=sl_off, c1=c4p, dst=None, dst_mode=3, dst_layout=1,
              out_name="score", out2_name="coord")
        dps = pl.buf_nhwc("dps", c3p, H2, W2)
        if self.upscale_method == "convtranspose":
            db = pl.buf_nhwc("db", _p32(4 * c3), H4, W4)
            pl.tc(P["desc.B"], da, _p32(4 * c3), c0_off=0, c0=c4p, dst=db)
            pl.tc(P["desc.up"], db, 4 * c3p, act=act, dst=dps, dst_mode=2)
        else:
            pl.tc(P["desc.B"], da, 4 * c3p, c0_off=0, c0=c4p, dst=dps, dst_mode=2)
        dA = pl.buf_nhwc("dA", c4p, H2, W2)
        pl.tc(P["desc.Aa"], dps, c4p, act=act, src1=skip, dst=dA)
        pl.tc(P["desc.Bb"], dA, self.nfeatures, dst=None, dst_layout=1, dst_c_total=self.nfeatures, out_name="feat")

    def _plan_seg_out_tc(self, pl, s7, packed_last):
        pl.tc(packed_last, s7, self.nClasses, dst=None, dst_layout=1, dst_c_total=self.nClasses, out_name="seg")


class KP2DTinyV3(_KP2DTinyBase):
    """Decoder-fusion model (kp2dtiny.py:650-1015): one score+loc head, descriptors from the seg trunk."""

    version = 3

    def __init__(self, use_color=True, do_cross=True, with_drop=True, nfeatures=256, device="cpu",
                 channel_dims=[32, 64, 128, 256, 256, 512], bn_momentum=0.1, nClasses=8, num_clusters=64,
                 downsample=3, use_attention=False, encoder_dim=None, mem_efficient=False,
                 upscale_method="pixelshuffle", remove_netvlad=False, leaky_relu=True, remove_softmax=False,
                 depth=False, global_descriptor_method="netvlad", **kwargs):
        super().__init__()
        if not use_color:
            raise NotImplementedError("use_color=False (1-channel input) is not on the hot path")
        self.device = device
        self.with_drop = with_drop
        self.nfeatures, self.downsample, self.nClasses = nfeatures, downsample, nClasses
        self.sample_segmentation = False
        self.use_color, self.do_cross, self.fuse_score_loc = use_color, do_cross, True
        self.remove_softmax, self.depth = remove_softmax, depth
        self.use_attention, self.leaky_relu = use_attention, leaky_relu
        self.remove_netvlad, self.upscale_method = remove_netvlad, upscale_method
        self.num_clusters, self.global_descriptor_method = num_clusters, global_descriptor_method
        self.mem_efficient = mem_efficient
        self.bn_momentum = bn_momentum
        self.channel_dims = list(channel_dims)
        c1, c2, c3, c4, c5, d1 = channel_dims
        self.encoder_dim = encoder_dim if encoder_dim is not None else c4
        self.backbone = _BackBone(3, c1, c2, c3, c4, 0.1)
        self.score_loc_head = _TaskHead(c4, c4, 3, bn_momentum)
        self.seg_head = _SegHead(c4, c5, c4 + c3, nClasses, d1, bn_momentum, use_attention, n_feat=nfeatures,
                                 depth=depth, upscale_method=upscale_method)
        self.vlad_head = _VPRHead(c4, self.encoder_dim, num_clusters, bn_momentum, remove_netvlad,
                                  global_descriptor_method)
        self._finish_init()

    def _pack_heads(self, P):
        h = self.score_loc_head
        tc = self.conv_backend == "tc"
        if self._merge_head_convs():
            P["sv.a"] = self._pk_block_pair(h.convDa, self.vlad_head.convlad1)
        else:
            P["sl.a"] = self._pk_block(h.convDa, tc=tc)
        # convDb (c4 -> 3) is launched as two tiny convs so that score (ch 0, sigmoid) and shift (ch 1:3, tanh)
        # land directly in their own output tensors (kp2dtiny.py:927-935) without a slicing copy.
        if tc:
            P["kp.b"] = ops.pack_head_pair_tc(h.convDb.weight, h.convDb.bias, cpad=_p32(self.channel_dims[3]),
                                              math=self.conv_math)
        else:
            P["sl.score"] = ops.pack_conv(h.convDb.weight[0:1], bias=h.convDb.bias[0:1])
            P["sl.shift"] = ops.pack_conv(h.convDb.weight[1:3], bias=h.convDb.bias[1:3])
        P["featB"] = self._pk_conv(self.seg_head.featB, tc=tc)
        if self.depth:
            P["featD"] = self._pk_conv(self.seg_head.featD, tc=tc)

    def _plan_heads(self, pl, xb, skip, act):
        P = self._packed
        c4 = self.channel_dims[3]
        B, _, H4, W4 = xb.shape
        sl = pl.buf("sl", c4, H4, W4)
        pl.conv(P["sl.a"], xb, c4, act=act, dst=sl)
        pl.conv(P["sl.score"], sl, 1, act=ops.ACT_SIGMOID, dst=torch.empty(B, 1, H4, W4, device=pl.device),
                out_name="score")
        pl.conv(P["sl.shift"], sl, 2, act=ops.ACT_TANH, dst=torch.empty(B, 2, H4, W4, device=pl.device),
                out_name="coord")

    def _plan_seg_out(self, pl, s7, packed_last):
        # segmentation.py:337-347 / :609-619: feat from the first half of the trunk, seg from the last half
        B, c5, H2, W2 = s7.shape  # c5 = width of the trunk's last block: 2 slices, 3 with depth
        ds = self.seg_head.dim_split
        pl.conv(self._packed["featB"], s7, self.nfeatures, c0_off=0, c0=ds,
                dst=torch.empty(B, self.nfeatures, H2, W2, device=pl.device), out_name="feat")
        if self.depth:  # middle slice -> featD -> sigmoid (segmentation.py:339-341, kp2dtiny.py:955-956)
            pl.out_shapes["depth"] = (B, 1, H2, W2)
            pl.conv(self._packed["featD"], s7, 1, c0_off=ds, c0=ds, act=ops.ACT_SIGMOID,
                    dst=torch.empty(B, 1, H2, W2, device=pl.device), out_name="depth")
        if self.remove_softmax:
            pl.conv(packed_last, s7, self.nClasses, c0_off=c5 - ds, c0=ds,
                    dst=torch.empty(B, self.nClasses, H2, W2, device=pl.device), out_name="seg")
        else:
            logits = pl.buf("seg_logits", self.nClasses, H2, W2)
            pl.conv(packed_last, s7, self.nClasses, c0_off=c5 - ds, c0=ds, dst=logits)
            pl.call(lambda outs: ops.softmax_channels(logits, out=outs["seg"]))  # Softmax2d (:942-943)

    def _plan_heads_tc(self, pl, xb, skip, act):
        P = self._packed
        c4p = _p32(self.channel_dims[3])
        B, H4, W4, _ = xb.shape
        if self._merge_head_convs():  # score_loc_head.convDa | vlad_head.convlad1 as one conv of the backbone output
            sva = pl.buf_nhwc("sva", 2 * c4p, H4, W4)
            pl.tc(P["sv.a"], xb, 2 * c4p, act=act, dst=sva)
            pl.bufs["v1_in_dva"] = sva
            sl = sva
        else:
            sl = pl.buf_nhwc("sl", c4p, H4, W4)
            pl.tc(P["sl.a"], xb, c4p, act=act, dst=sl)
        pl.tc(P["kp.b"], sl, 3, c0_off=0, c0=c4p, dst=None, dst_mode=3, dst_layout=1, out_name="score",
              out2_name="coord")

    def _plan_seg_out_tc(self, pl, s7, packed_last):
        B, H2, W2, _ = s7.shape
        c5 = self.seg_head.c_last      # real width of the trunk's last block: 2 slices, 3 with depth
        ds = self.seg_head.dim_split   # real channels per slice; each conv reads a 32-channel (padded) window
        dsp = _p32(ds)
        pl.tc(self._packed["featB"], s7, self.nfeatures, c0_off=0, c0=dsp, dst=None, dst_layout=1,
              dst_c_total=self.nfeatures, out_name="feat")
        if self.depth:
            pl.out_shapes["depth"] = (B, 1, H2, W2)
            pl.tc(self._packed["featD"], s7, 1, c0_off=ds, c0=dsp, act=ops.ACT_SIGMOID, dst=None, dst_layout=1,
                  dst_c_total=1, out_name="depth")
        if self.remove_softmax:
            pl.tc(packed_last, s7, self.nClasses, c0_off=c5 - ds, c0=dsp, dst=None, dst_layout=1,
                  dst_c_total=self.nClasses, out_name="seg")
        else:
            logits = pl.buf("seg_logits", self.nClasses, H2, W2)
            pl.tc(packed_last, s7, self.nClasses, c0_off=c5 - ds, c0=dsp, dst=logits, dst_layout=1)
            pl.call(lambda outs: ops.softmax_channels(logits, out=outs["seg"]))
