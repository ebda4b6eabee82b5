"""Functional torch-CPU restatement of KP2DTinyV2 / KP2DTinyV3 (TEST INFRASTRUCTURE).

Works on a ``state_dict`` with the reference's key names.  All citations are
``path:line`` inside the upstream repository (``src/kp2dtiny/...``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass(frozen=True)
class Arch:
    """Architecture numbers taken from the config letters (models/kp2dtiny.py:46-218)."""

    version: int  # 2 = dedicated decoders, 3 = decoder fusion
    channel_dims: tuple  # c1..c5, d1
    nfeatures: int = 32
    n_classes: int = 28
    encoder_dim: int = 64
    num_clusters: int = 64
    use_attention: bool = False
    leaky_relu: bool = True
    downsample: int = 2
    global_descriptor_method: str = "netvlad"  # "gem" / "convap": letters GEM_*, CONVAP_* (kp2dtiny.py:64-82, 135-144)
    depth: bool = False  # constructor kwarg depth=True (kp2dtiny.py:315,402-437; segmentation.py:189-191,284-287)
    upscale_method: str = "pixelshuffle"  # "convtranspose" with to_mcu=True, which also sets leaky_relu=False (:271-274)

    @property
    def cell(self) -> int:
        return 2 ** self.downsample  # kp2dtiny.py:455


def arch_for(letter: str, v3: bool, n_classes: int, depth: bool = False, to_mcu: bool = False) -> Arch:
    import dataclasses

    a = dataclasses.replace(_arch_for(letter, v3, n_classes), depth=depth)
    if to_mcu:  # get_config(to_mcu=True), kp2dtiny.py:271-274
        a = dataclasses.replace(a, upscale_method="convtranspose", leaky_relu=False)
    return a


def upsample(x: Tensor, sd: SD, p: str, a: Arch) -> Tensor:
    """PixelShuffle(2), or TransposedConvUpsampleModel (modules/base.py:80-117): ConvTranspose2d(c, c/4, 3, stride 2,
    padding 1, output_padding 1, no bias) -> BN(eval) -> LeakyReLU | ReLU."""
    if a.upscale_method == "pixelshuffle":
        return F.pixel_shuffle(x, 2)
    y = F.conv_transpose2d(x, sd[p + ".transposed_conv.weight"], None, stride=2, padding=1, output_padding=1)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], training=False, eps=1e-5)
    return F.leaky_relu(y, 0.01) if a.leaky_relu else F.relu(y)


def _arch_for(letter: str, v3: bool, n_classes: int) -> Arch:
    """Letter -> numbers; mirrors TINY_S / TINY_N / V3_S / V3_N ... (kp2dtiny.py:46-166)."""
    method = "netvlad"
    for prefix, m in (("GEM_", "gem"), ("CONVAP_", "convap")):
        if letter.startswith(prefix):
            letter, method = letter[len(prefix):], m
    small = letter.startswith("S")
    att = letter.endswith("_A")
    if letter in ("D", "D_A"):  # LARGE_D / LARGE_D_V3 / LARGE_D_A_V3 (kp2dtiny.py:168-196): ConvAP, 128-d descriptors
        att = att or not v3     # V2 "D" has use_attention=True
        return Arch(3 if v3 else 2, (64, 128, 128, 256, 256, 512), 128, n_classes, 128, 64, att, True, 2, "convap")
    if letter == "F":  # TINY_F (kp2dtiny.py:112-118): cell 8, 64-d descriptors, encoder_dim defaults to c4, NetVLAD K = 64
        return Arch(2, (16, 32, 64, 128, 128, 256), 64, n_classes, 128, 64, False, True, 3, "netvlad")
    if small:
        dims, enc = (16, 32, 32, 64, 64, 128), 64
    else:
        dims, enc = (16, 24, 24, 48, 48, 96), 48
    # V2 'N'/'N_A' set num_clusters=32 (kp2dtiny.py:84-102); every other letter keeps the
    # constructor default of 64 (kp2dtiny.py:308, :690).
    k = 32 if (not small and not v3) else 64
    return Arch(3 if v3 else 2, dims, 32, n_classes, enc, k, att, True, 2, method)


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def conv_bn_act(x: Tensor, sd: SD, p: str, leaky: bool = True) -> Tensor:
    """modules/base.py:39-46  conv3x3(no bias) -> BN(eval) -> LeakyReLU(0.01)|ReLU."""
    y = F.conv2d(x, sd[p + ".conv.weight"], None, stride=1, padding=1)
    y = F.batch_norm(
        y,
        sd[p + ".bn.running_mean"],
        sd[p + ".bn.running_var"],
        sd[p + ".bn.weight"],
        sd[p + ".bn.bias"],
        training=False,
        eps=1e-5,
    )
    return F.leaky_relu(y, 0.01) if leaky else F.relu(y)


def conv_bias(x: Tensor, sd: SD, p: str) -> Tensor:
    """Plain biased 3x3 conv (heads.py:33,96,102; segmentation.py:154,339-342)."""
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=1, padding=1)


def backbone(x: Tensor, sd: SD, a: Arch):
    """modules/encoders.py:105-129 (downsample=2: pools after conv1b and conv3b)."""
    lk = a.leaky_relu
    x = conv_bn_act(x, sd, "backbone.conv1a", lk)
    x = conv_bn_act(x, sd, "backbone.conv1b", lk)
    if a.downsample >= 2:
        x = F.max_pool2d(x, 2, 2)
    x = conv_bn_act(x, sd, "backbone.conv2a", lk)
    x = conv_bn_act(x, sd, "backbone.conv2b", lk)
    if a.downsample >= 3:
        x = F.max_pool2d(x, 2, 2)
    x = conv_bn_act(x, sd, "backbone.conv3a", lk)
    skip = conv_bn_act(x, sd, "backbone.conv3b", lk)
    x = F.max_pool2d(skip, 2, 2) if a.downsample >= 1 else skip
    x = conv_bn_act(x, sd, "backbone.conv4a", lk)
    x = conv_bn_act(x, sd, "backbone.conv4b", lk)
    return x, skip


def simple_task_head(x: Tensor, sd: SD, p: str, a: Arch) -> Tensor:
    """modules/decoders/heads.py:28-35."""
    return conv_bias(conv_bn_act(x, sd, p + ".convDa", a.leaky_relu), sd, p + ".convDb")


def upscale_head(x: Tensor, skip: Tensor, sd: SD, a: Arch) -> Tensor:
    """modules/decoders/heads.py:91-104 (pixelshuffle variant)."""
    p = "desc_head"
    y = conv_bn_act(x, sd, p + ".convA", a.leaky_relu)
    y = conv_bias(y, sd, p + ".convB")
    y = upsample(y, sd, p + ".upsample", a)
    y = torch.cat([y, skip], dim=1)
    y = conv_bn_act(y, sd, p + ".confAa", a.leaky_relu)
    return conv_bias(y, sd, p + ".confBb")


def channel_layernorm(x: Tensor, g: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """modules/segformer.py:70-73 -- note (std + eps), biased variance."""
    std = torch.var(x, dim=1, unbiased=False, keepdim=True).sqrt()
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) / (std + eps) * g + b


def efficient_self_attention(x: Tensor, sd: SD, p: str, heads: int = 4) -> Tensor:
    """modules/segformer.py:100-138: 1x1 q, 2x2/s2 kv, per-head softmax(q k^T d^-1/2) v, 1x1 out."""
    B, C, h, w = x.shape
    d = C // heads
    q = F.conv2d(x, sd[p + ".to_q.weight"])
    kv = F.conv2d(x, sd[p + ".to_kv.weight"], stride=2)
    k, v = kv[:, :C], kv[:, C:]

    def split(t):  # b (h c) x y -> (b h) (x y) c      segformer.py:22-30
        b, _, hh, ww = t.shape
        return t.reshape(b, heads, d, hh * ww).permute(0, 1, 3, 2).reshape(b * heads, hh * ww, d)

    q, k, v = split(q), split(k), split(v)
    # softmax is per query row, so the rows can be processed in blocks: the same arithmetic as the reference's
    # single matmul, without its (b h) x n x m temporary (4.3 GB per 512x1024 frame)
    rows = max(1, (1 << 26) // max(1, k.shape[1]))
    out = torch.empty_like(q)
    for r0 in range(0, q.shape[1], rows):
        sim = torch.matmul(q[:, r0:r0 + rows], k.transpose(-2, -1)) * (d ** -0.5)
        out[:, r0:r0 + rows] = torch.matmul(sim.softmax(dim=-1), v)  # (b h) n d
    out = out.reshape(B, heads, h * w, d).permute(0, 1, 3, 2).reshape(B, C, h, w)
    return F.conv2d(out, sd[p + ".to_out.weight"])


def mix_ffn(x: Tensor, sd: SD, p: str) -> Tensor:
    """modules/segformer.py:193-205: 1x1 -> dw3x3 -> 1x1 -> GELU(erf) -> 1x1, all biased."""
    C2 = sd[p + ".net.0.weight"].shape[0]
    y = F.conv2d(x, sd[p + ".net.0.weight"], sd[p + ".net.0.bias"])
    y = F.conv2d(y, sd[p + ".net.1.net.0.weight"], sd[p + ".net.1.net.0.bias"], padding=1, groups=C2)
    y = F.conv2d(y, sd[p + ".net.1.net.1.weight"], sd[p + ".net.1.net.1.bias"])
    y = F.gelu(y)
    return F.conv2d(y, sd[p + ".net.3.weight"], sd[p + ".net.3.bias"])


def segformer_attention_module(x: Tensor, sd: SD, p: str) -> Tensor:
    """modules/segformer.py:209-220: PreNorm(attention) then PreNorm(MixFFN); no residual adds."""
    x = efficient_self_attention(channel_layernorm(x, sd[p + ".att.norm.g"], sd[p + ".att.norm.b"]), sd, p + ".att.fn")
    x = mix_ffn(channel_layernorm(x, sd[p + ".mff.norm.g"], sd[p + ".mff.norm.b"]), sd, p + ".mff.fn")
    return x


def seg_trunk(x: Tensor, skip: Tensor, sd: SD, a: Arch, head: str = "seg_head") -> Tensor:
    """Shared trunk of the four segmentation heads up to (and including) the last conv block.

    plain:     segmentation.py:126-152 (V2) / :314-334 (V3)   -> convs[0..7]
    attention: segmentation.py:442-463 (V2) / :588-608 (V3)   -> convs[0..6]
    """
    lk, p = a.leaky_relu, head + ".convs."
    s = conv_bn_act(x, sd, p + "0", lk)
    if a.use_attention:
        s = segformer_attention_module(s, sd, p + "1")
        s = F.max_pool2d(s, 2, 2)
        s = segformer_attention_module(s, sd, p + "2")
        nxt = 3
    else:
        s = conv_bn_act(s, sd, p + "1", lk)
        s = F.max_pool2d(s, 2, 2)
        s = conv_bn_act(s, sd, p + "2", lk)
        s = conv_bn_act(s, sd, p + "3", lk)
        nxt = 4
    s = conv_bn_act(s, sd, p + str(nxt), lk)  # -> d1
    s = torch.cat([upsample(s, sd, head + ".upsample", a), x], dim=1)
    s = conv_bn_act(s, sd, p + str(nxt + 1), lk)
    s = conv_bn_act(s, sd, p + str(nxt + 2), lk)  # -> d1
    s = torch.cat([upsample(s, sd, head + ".upsample2", a), skip], dim=1)
    return conv_bn_act(s, sd, p + str(nxt + 3), lk)


def netvlad(x: Tensor, sd: SD, p: str = "vlad_head.netvlad") -> Tensor:
    """modules/aggregators/netvlad.py:79-106 written as V = A X^T - c * sum(A)."""
    N, C = x.shape[:2]
    cent = sd[p + ".centroids"]  # (K, C)
    K = cent.shape[0]
    xn = F.normalize(x, p=2.0, dim=1)
    soft = F.conv2d(xn, sd[p + ".conv.weight"]).view(N, K, -1).softmax(dim=1)  # (N,K,S)
    xf = xn.view(N, C, -1)  # (N,C,S)
    vlad = torch.matmul(soft, xf.transpose(1, 2)) - cent.unsqueeze(0) * soft.sum(dim=2, keepdim=True)
    vlad = F.normalize(vlad, p=2.0, dim=2)
    return F.normalize(vlad.reshape(N, -1), p=2.0, dim=1)


def netvlad_literal(x: Tensor, sd: SD, p: str = "vlad_head.netvlad") -> Tensor:
    """Same layer, with the residual tensor materialised exactly like netvlad.py:95-101."""
    N, C = x.shape[:2]
    cent = sd[p + ".centroids"]
    K = cent.shape[0]
    xn = F.normalize(x, p=2.0, dim=1)
    soft = F.conv2d(xn, sd[p + ".conv.weight"]).view(N, K, -1).softmax(dim=1)
    xf = xn.view(N, C, -1)
    resid = xf.unsqueeze(1) - cent.view(1, K, C, 1)  # (N,K,C,S)
    vlad = (resid * soft.unsqueeze(2)).sum(dim=-1)
    vlad = F.normalize(vlad, p=2.0, dim=2)
    return F.normalize(vlad.reshape(N, -1), p=2.0, dim=1)


def gem(x: Tensor, sd: SD, p: str = "vlad_head.netvlad", eps: float = 1e-6) -> Tensor:
    """modules/aggregators/gem.py:21-33 with unshuffle=4 (vpr.py:71): PixelUnshuffle(4), then
    (avg over the whole map of clamp(x, eps)^p)^(1/p), flattened; no normalisation."""
    pw = sd[p + ".p"]
    x = F.pixel_unshuffle(x, 4)
    return F.avg_pool2d(x.clamp(min=eps).pow(pw), (x.size(-2), x.size(-1))).pow(1.0 / pw).flatten(1)


def convap(x: Tensor, sd: SD, p: str = "vlad_head.netvlad", s: int = 4) -> Tensor:
    """modules/aggregators/convap.py:29-37 with s1 = s2 = 4 (vpr.py:74-75): biased 1x1 conv,
    AdaptiveAvgPool2d((4,4)), flatten, L2 normalise."""
    y = F.conv2d(x, sd[p + ".channel_pool.weight"], sd[p + ".channel_pool.bias"])
    y = F.adaptive_avg_pool2d(y, (s, s))
    return F.normalize(y.flatten(1), p=2.0, dim=1)


def vpr_head(x: Tensor, sd: SD, a: Arch, only_encoder: bool = False) -> Tensor:
    """modules/decoders/vpr.py:78-89."""
    v = conv_bn_act(x, sd, "vlad_head.convlad1", a.leaky_relu)
    v = conv_bn_act(v, sd, "vlad_head.convlad2", a.leaky_relu)
    v = conv_bn_act(v, sd, "vlad_head.convlad3", a.leaky_relu)
    if only_encoder:
        return F.normalize(v, p=2.0, dim=1)
    if a.global_descriptor_method == "gem":
        return gem(v, sd)
    if a.global_descriptor_method == "convap":
        return convap(v, sd)
    return netvlad(v, sd)


# ----------------------------------------------------------------------------------------------
# whole-model forward + post_processing
# ----------------------------------------------------------------------------------------------
@torch.no_grad()
def forward(x: Tensor, sd: SD, a: Arch) -> Dict[str, Tensor]:
    """KP2DTinyV2.forward (kp2dtiny.py:552-591) / KP2DTinyV3.forward (:906-957), inference mode
    (``training is False``): returns raw tanh shift under 'coord', dense 'feat', seg logits (V2) or
    Softmax2d probabilities (V3, :942-943)."""
    xb, skip = backbone(x, sd, a)
    if a.version == 2:
        score = simple_task_head(xb, sd, "score_head", a).sigmoid()
        shift = simple_task_head(xb, sd, "loc_head", a).tanh()
        feat = upscale_head(xb, skip, sd, a)
        last = "seg_head.convs.7" if a.use_attention else "seg_head.convs.8"
        seg = conv_bias(seg_trunk(xb, skip, sd, a), sd, last)
        if a.depth:  # a second segmentation head with one output channel, then sigmoid (kp2dtiny.py:402-437,588-590)
            depth = conv_bias(seg_trunk(xb, skip, sd, a, "depth_head"), sd, last.replace("seg_head", "depth_head")).sigmoid()
    else:
        sl = simple_task_head(xb, sd, "score_loc_head", a)
        score, shift = sl[:, 0:1].sigmoid(), sl[:, 1:3].tanh()
        t = seg_trunk(xb, skip, sd, a)
        ds = a.channel_dims[4] // 2  # dim_split = c_hidden // 2   segmentation.py:187,493
        feat = conv_bias(t[:, :ds], sd, "seg_head.featB")
        last = "seg_head.convs.7" if a.use_attention else "seg_head.convs.8"
        seg = conv_bias(t[:, -ds:], sd, last)
        seg = seg.softmax(dim=1)  # Softmax2d
        if a.depth:  # the trunk's last block is 3*ds wide; the middle slice feeds featD (segmentation.py:339-341), :955-956
            depth = F.conv2d(t[:, ds:2 * ds], sd["seg_head.featD.weight"], None, padding=1).sigmoid()
    vlad = vpr_head(xb, sd, a)
    out = {"score": score, "coord": shift, "feat": feat, "vlad": vlad, "seg": seg}
    if a.depth:
        out["depth"] = depth
    return out


@torch.no_grad()
def post_processing(out: Dict[str, Tensor], H: int, W: int, a: Arch,
                    sample_segmentation: bool = False) -> Dict[str, Tensor]:
    """KP2DTinyV*.post_processing with ``training is False`` (kp2dtiny.py:593-647 / :959-1015)."""
    score, shift, feat, seg = out["score"], out["coord"], out["feat"], out["seg"]
    B, _, Hc, Wc = score.shape
    # remove_border (:520-528)
    mask = torch.ones(B, 1, Hc, Wc, dtype=score.dtype)
    mask[:, :, 0] = 0
    mask[:, :, Hc - 1] = 0
    mask[:, :, :, 0] = 0
    mask[:, :, :, Wc - 1] = 0
    score = score * mask
    # coordinates (:597-614); cross_ratio = 2.0 (:339)
    step = (a.cell - 1) / 2.0
    ys, xs = torch.meshgrid(torch.linspace(0, Hc - 1, Hc), torch.linspace(0, Wc - 1, Wc), indexing="ij")
    base = torch.stack([xs, ys], 0).unsqueeze(0).repeat(B, 1, 1, 1).mul(a.cell) + step
    coord = base + shift * (2.0 * step)
    coord = torch.stack([coord[:, 0].clamp(0, W - 1), coord[:, 1].clamp(0, H - 1)], 1)
    # normalize_coord (:642-647) + sample_feat (:627-631)
    cn = torch.stack([coord[:, 0] / (float(W - 1) / 2.0) - 1.0, coord[:, 1] / (float(H - 1) / 2.0) - 1.0], 1)
    cn = cn.permute(0, 2, 3, 1)
    fs = F.grid_sample(feat, cn, mode="bilinear", padding_mode="zeros", align_corners=True)
    fs = fs / torch.norm(fs, p=2, dim=1, keepdim=True)
    # sample_seg (:633-640 V2 softmax+argmax; :1001-1008 V3 argmax of the probabilities)
    if sample_segmentation:
        seg = F.grid_sample(seg, cn, mode="nearest", padding_mode="zeros", align_corners=True)
    if a.version == 2:
        seg = seg.softmax(dim=1)
    seg = seg.argmax(1).unsqueeze(1)
    return {"score": score, "coord": coord, "feat": fs, "vlad": out["vlad"], "seg": seg}


# ----------------------------------------------------------------------------------------------
# algorithmic work; used by bench.py for roofline figures
# ----------------------------------------------------------------------------------------------
def conv_flops_per_frame(a: Arch, H: int, W: int) -> float:
    """2*MACs of every conv on the forward path for one frame (dense 3x3 unless noted)."""
    c1, c2, c3, c4, c5, d1 = a.channel_dims
    h2, w2 = H // 2, W // 2
    h4, w4 = h2 // 2, w2 // 2
    h8, w8 = h4 // 2, w4 // 2
    f = 0.0

    def c3x3(ci, co, h, w):
        return 2.0 * 9 * ci * co * h * w

    f += c3x3(3, c1, H, W) + c3x3(c1, c2, H, W)
    f += c3x3(c2, c2, h2, w2) + c3x3(c2, c3, h2, w2) + c3x3(c3, c3, h2, w2) + c3x3(c3, c4, h2, w2)
    f += 2 * c3x3(c4, c4, h4, w4)
    if a.version == 2:
        f += 2 * c3x3(c4, c4, h4, w4) + c3x3(c4, 1, h4, w4) + c3x3(c4, 2, h4, w4)
        f += c3x3(c4, c4, h4, w4) + c3x3(c4, 4 * c3, h4, w4) + c3x3(c3 + c4, c4, h2, w2) + c3x3(c4, a.nfeatures, h2, w2)
    else:
        f += c3x3(c4, c4, h4, w4) + c3x3(c4, 3, h4, w4)
    # seg trunk
    f += c3x3(c4, c5, h4, w4)
    if a.use_attention:
        for (h, w) in ((h4, w4), (h8, w8)):
            n, nk, C = h * w, (h // 2) * (w // 2), c5
            f += 2.0 * C * C * n + 2.0 * 4 * C * 2 * C * nk + 2.0 * 2 * n * nk * C + 2.0 * C * C * n
            f += 2.0 * C * 2 * C * n + 2.0 * 9 * 2 * C * n + 2.0 * 2 * C * 2 * C * n + 2.0 * 2 * C * C * n
    else:
        f += c3x3(c5, c5, h4, w4) + 2 * c3x3(c5, c5, h8, w8)
    f += c3x3(c5, d1, h8, w8) + c3x3(c5 + d1 // 4, c5, h4, w4) + c3x3(c5, d1, h4, w4) + c3x3(c3 + c4, c5, h2, w2)
    if a.version == 2:
        f += c3x3(c5, a.n_classes, h2, w2)
    else:
        f += c3x3(c5 // 2, a.nfeatures, h2, w2) + c3x3(c5 // 2, a.n_classes, h2, w2)
    # vlad
    e = a.encoder_dim
    f += c3x3(c4, e, h4, w4) + 2 * c3x3(e, e, h4, w4) + 2.0 * e * a.num_clusters * h4 * w4 * 2
    return f


def algorithmic_bytes_per_frame(a: Arch, H: int, W: int, with_decode: bool = True) -> float:
    """fp32 input + forward-dict outputs (+ decode outputs) for one frame, SURVEY.md §8(d)."""
    h2, w2, h4, w4 = H // 2, W // 2, H // 4, W // 4
    b = 3 * H * W + 3 * h4 * w4 + a.nfeatures * h2 * w2 + a.n_classes * h2 * w2 + a.encoder_dim * a.num_clusters
    if with_decode:
        b += (3 + a.nfeatures) * h4 * w4 + 2 * h2 * w2  # score,coord,feat f32 + seg int64
    return 4.0 * b
