"""``torch.ops.nanovs.*`` -- the PyTorch custom-op face of libnanovs.so (last row).

The reference has no operator interface of its own: its callers use ``KP2DTinyV2/V3.forward`` and ``post_processing``
(src/kp2dtiny/models/kp2dtiny.py:552-647, :906-1015), ``KP2DtinyFrontend.run`` (src/visual_odometry/frontend.py:78-129),
``BfFeatureMatcher.match`` (feature_matcher.py:89-98), ``faiss.IndexFlatL2`` (evaluation/global_descriptor.py:55-60) and
``estimatePose`` (visual_odometry.py:383-412).  Every one of those statements is served by one operator registered
here, and the module surface in this package calls nothing else:

    kp2dtiny_forward   whole-model launch plan (backbone + heads + NetVLAD)         kp2dtiny.py:552-591 / :906-957
    decode, seg_argmax post_processing                                              kp2dtiny.py:593-647 / :959-1015
    select_keypoints   threshold + semantic filter + top-k                          frontend.py:94-126
    match, match_batch 2-NN + ratio + one-to-one / mutual NN                        feature_matcher.py:89-98,179-209
    pose_batch         findEssentialMat + recoverPose                               visual_odometry.py:383-412
    flat_l2_prepare / flat_l2_search / topk_merge                                   global_descriptor.py:55-60
    conv_tc, conv, attention, netvlad, channel_layernorm, dwconv3x3, softmax_channels, preprocess_u8
                       the building blocks of the plan, exposed for direct use      modules/*.py

Registration is for the CUDA dispatch key ONLY: a CPU tensor raises ``NotImplementedError`` from the dispatcher --
there is no CPU kernel and no fallback by design.  Every operator has a fake ("meta") implementation, so shapes and
dtypes propagate under FakeTensorMode / torch.export without touching the GPU.  The CUDA implementations hand raw
device pointers and the current stream to the C ABI (include/nanovs.h) through ``ops``/``_cabi`` (ctypes).
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence, Tuple

import torch

from . import ops

_LIB = torch.library.Library("nanovs", "DEF")
_MODELS: "weakref.WeakValueDictionary[int, torch.nn.Module]" = weakref.WeakValueDictionary()
_INDEXES: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
OP_NAMES: List[str] = []


def _define(schema: str, cuda_impl, fake_impl) -> None:
    name = schema.split("(", 1)[0]
    _LIB.define(schema)
    _LIB.impl(name, cuda_impl, "CUDA")
    torch.library.register_fake("nanovs::" + name, fake_impl, lib=_LIB)
    OP_NAMES.append(name)


def register_model(model: torch.nn.Module) -> int:
    """Handle under which ``kp2dtiny_forward`` finds the module's packed weights and launch plans."""
    h = id(model)
    _MODELS[h] = model
    return h


def _model(handle: int):
    m = _MODELS.get(int(handle))
    if m is None:
        raise RuntimeError(f"nanovs::kp2dtiny_forward: unknown model handle {handle}")
    return m


# ------------------------------------------------------------------------------------------------------------------
# whole-model forward
# ------------------------------------------------------------------------------------------------------------------
def _fwd_cuda(x: torch.Tensor, handle: int, unit_input: bool) -> List[torch.Tensor]:
    m = _model(handle)
    out = m._forward_impl(x, unit_input)
    return [out[k] for k in m._forward_keys()]


def _fwd_fake(x: torch.Tensor, handle: int, unit_input: bool) -> List[torch.Tensor]:
    m = _model(handle)
    if x.dtype == torch.uint8:
        B, H, W = x.shape[0], x.shape[1], x.shape[2]
    else:
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
    return [x.new_empty(s, dtype=torch.float32) for s in m._forward_shapes(B, H, W)]


_define("kp2dtiny_forward(Tensor x, int model, bool unit_input) -> Tensor[]", _fwd_cuda, _fwd_fake)


# ------------------------------------------------------------------------------------------------------------------
# decode / selection
# ------------------------------------------------------------------------------------------------------------------
def _decode_cuda(score, shift, feat, H: int, W: int, cell: int, cross_ratio: float):
    o_s, o_c, o_f = ops.decode(score, shift, feat, H, W, cell, cross_ratio)
    return o_s, o_c, (o_f if o_f is not None else score.new_empty(0))


def _decode_fake(score, shift, feat, H: int, W: int, cell: int, cross_ratio: float):
    B, _, Hc, Wc = score.shape
    o_f = score.new_empty(B, feat.shape[1], Hc, Wc) if feat is not None else score.new_empty(0)
    return torch.empty_like(score), torch.empty_like(shift), o_f


_define("decode(Tensor score, Tensor shift, Tensor? feat, int H, int W, int cell, float cross_ratio) "
        "-> (Tensor, Tensor, Tensor)", _decode_cuda, _decode_fake)


def _argmax_cuda(seg, coord, H: int, W: int):
    return ops.seg_argmax(seg, coord, H, W)


def _argmax_fake(seg, coord, H: int, W: int):
    B = seg.shape[0]
    hw = coord.shape[2:] if coord is not None else seg.shape[2:]
    return seg.new_empty(B, 1, hw[0], hw[1], dtype=torch.int64)


_define("seg_argmax(Tensor seg, Tensor? coord, int H, int W) -> Tensor", _argmax_cuda, _argmax_fake)


def _select_k(score, top_k: int) -> int:
    n_cells = score.shape[1] * score.shape[2] * score.shape[3]
    return n_cells if top_k <= 0 else min(top_k, n_cells)


def _select_cuda(score, coord, feat, seg_cells, classes: Sequence[int], thresh: float, top_k: int):
    r = ops.select_keypoints(score, coord, feat, thresh, top_k, seg_cells=seg_cells,
                             classes_to_filter=list(classes) if classes else None)
    e = score.new_empty(0)
    return (r["pts"], r["desc"] if r["desc"] is not None else e, r["score"], r["cell"],
            r["label"] if r["label"] is not None else e.to(torch.int64), r["count"])


def _select_fake(score, coord, feat, seg_cells, classes: Sequence[int], thresh: float, top_k: int):
    B, k = score.shape[0], _select_k(score, top_k)
    e = score.new_empty(0)
    return (score.new_empty(B, k, 2), score.new_empty(B, k, feat.shape[1]) if feat is not None else e,
            score.new_empty(B, k), score.new_empty(B, k, dtype=torch.int32),
            score.new_empty(B, k, dtype=torch.int64) if seg_cells is not None else e.to(torch.int64),
            score.new_empty(B, dtype=torch.int32))


_define("select_keypoints(Tensor score, Tensor coord, Tensor? feat, Tensor? seg_cells, int[] classes_to_filter, "
        "float thresh, int top_k) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", _select_cuda, _select_fake)


def select_keypoints(score, coord, feat, thresh: float, top_k: int, seg_cells=None, classes_to_filter=None) -> dict:
    """``torch.ops.nanovs.select_keypoints`` with the dict return of ``ops.select_keypoints``."""
    pts, desc, sc, cell, label, count = torch.ops.nanovs.select_keypoints(
        score, coord, feat, seg_cells, list(classes_to_filter) if classes_to_filter else [], float(thresh), int(top_k))
    return {"pts": pts, "desc": desc if feat is not None else None, "score": sc, "cell": cell,
            "label": label if seg_cells is not None else None, "count": count}


# ------------------------------------------------------------------------------------------------------------------
# matching / pose
# ------------------------------------------------------------------------------------------------------------------
def _match_cuda(des1, des2, ratio: float, mode: int):
    r = ops.match(des1, des2, ratio=ratio, mode=mode)
    if mode == 2:
        idx, dist = r
        e = des1.new_empty(0, dtype=torch.int32)
        return idx, e, dist, e
    return r


def _match_fake(des1, des2, ratio: float, mode: int):
    n1 = des1.shape[0]
    i32 = dict(dtype=torch.int32)
    if mode == 2:
        return des1.new_empty(n1, 2, **i32), des1.new_empty(0, **i32), des1.new_empty(n1, 2), des1.new_empty(0, **i32)
    return des1.new_empty(n1, **i32), des1.new_empty(n1, **i32), des1.new_empty(n1), des1.new_empty(1, **i32)


_define("match(Tensor des1, Tensor des2, float ratio, int mode) -> (Tensor, Tensor, Tensor, Tensor)",
        _match_cuda, _match_fake)


def _match_batch_cuda(desc, counts, pair_a, pair_b, ratio: float, mode: int):
    return ops.match_batch(desc, counts, pair_a, pair_b, ratio=ratio, mode=mode)


def _match_batch_fake(desc, counts, pair_a, pair_b, ratio: float, mode: int):
    P, kmax = pair_a.shape[0], desc.shape[1]
    i32 = dict(dtype=torch.int32)
    return desc.new_empty(P, kmax, **i32), desc.new_empty(P, kmax, **i32), desc.new_empty(P, kmax), desc.new_empty(P, **i32)


_define("match_batch(Tensor desc, Tensor counts, Tensor pair_a, Tensor pair_b, float ratio, int mode) "
        "-> (Tensor, Tensor, Tensor, Tensor)", _match_batch_cuda, _match_batch_fake)


def _pose_cuda(pts, pair_a, pair_b, count, idx1, idx2, intrinsics: Sequence[float], threshold: float, iters: int,
               seed: int, refine: int):
    r = ops.pose_batch(pts, pair_a, pair_b, count, idx1, idx2, intrinsics=tuple(intrinsics), threshold=threshold,
                       iters=iters, seed=seed, refine=refine)
    return r["E"], r["R"], r["t"], r["mask"], r["inliers"]


def _pose_fake(pts, pair_a, pair_b, count, idx1, idx2, intrinsics: Sequence[float], threshold: float, iters: int,
               seed: int, refine: int):
    P, kmax = pair_a.shape[0], pts.shape[1]
    return (pts.new_empty(P, 3, 3), pts.new_empty(P, 3, 3), pts.new_empty(P, 3),
            pts.new_empty(P, kmax, dtype=torch.uint8), pts.new_empty(P, dtype=torch.int32))


_define("pose_batch(Tensor pts, Tensor pair_a, Tensor pair_b, Tensor count, Tensor? idx1, Tensor? idx2, "
        "float[] intrinsics, float threshold, int iters, int seed, int refine) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor)", _pose_cuda, _pose_fake)


def _pose_adaptive_cuda(pts, pair_a, pair_b, count, idx1, idx2, intrinsics: Sequence[float], threshold: float, iters: int,
                        seed: int, refine: int, confidence: float, round_size: int):
    r = ops.pose_batch(pts, pair_a, pair_b, count, idx1, idx2, intrinsics=tuple(intrinsics), threshold=threshold,
                       iters=iters, seed=seed, refine=refine, confidence=confidence, round_size=round_size)
    return r["E"], r["R"], r["t"], r["mask"], r["inliers"], r["iters"]


def _pose_adaptive_fake(pts, pair_a, pair_b, count, idx1, idx2, intrinsics: Sequence[float], threshold: float, iters: int,
                        seed: int, refine: int, confidence: float, round_size: int):
    P = pair_a.shape[0]
    return _pose_fake(pts, pair_a, pair_b, count, idx1, idx2, intrinsics, threshold, iters, seed, refine) + \
        (pts.new_empty(P, dtype=torch.int32),)


_define("pose_batch_adaptive(Tensor pts, Tensor pair_a, Tensor pair_b, Tensor count, Tensor? idx1, Tensor? idx2, "
        "float[] intrinsics, float threshold, int iters, int seed, int refine, float confidence, int round_size) "
        "-> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", _pose_adaptive_cuda, _pose_adaptive_fake)


def pose_batch(pts, pair_a, pair_b, count, idx1=None, idx2=None, intrinsics=(1.0, 1.0, 0.0, 0.0),
               threshold: float = 0.0003, iters: int = 512, seed: int = 0, refine: int = 0, confidence: float = 0.0,
               round_size: int = 64) -> dict:
    if confidence > 0.0:
        E, R, t, mask, inl, used = torch.ops.nanovs.pose_batch_adaptive(
            pts, pair_a, pair_b, count, idx1, idx2, [float(v) for v in intrinsics], float(threshold), int(iters),
            int(seed), int(refine), float(confidence), int(round_size))
        return {"E": E, "R": R, "t": t, "mask": mask, "inliers": inl, "iters": used}
    E, R, t, mask, inl = torch.ops.nanovs.pose_batch(pts, pair_a, pair_b, count, idx1, idx2,
                                                     [float(v) for v in intrinsics], float(threshold), int(iters),
                                                     int(seed), int(refine))
    return {"E": E, "R": R, "t": t, "mask": mask, "inliers": inl}


# ------------------------------------------------------------------------------------------------------------------
# retrieval
# ------------------------------------------------------------------------------------------------------------------
def register_index(index) -> int:
    h = id(index)
    _INDEXES[h] = index
    return h


def _search_cuda(q, index: int, k: int, id_offset: int):
    ix = _INDEXES.get(int(index))
    if ix is None:
        raise RuntimeError(f"nanovs::flat_l2_search: unknown index handle {index}")
    return ix._search_impl(q, k, id_offset)


def _search_fake(q, index: int, k: int, id_offset: int):
    return q.new_empty(q.shape[0], k), q.new_empty(q.shape[0], k, dtype=torch.int64)


_define("flat_l2_search(Tensor q, int index, int k, int id_offset) -> (Tensor, Tensor)", _search_cuda, _search_fake)


def _merge_cuda(D_parts, I_parts):
    from .retrieval import _merge_topk_impl

    return _merge_topk_impl(D_parts, I_parts)


def _merge_fake(D_parts, I_parts):
    _, nq, k = D_parts.shape
    return D_parts.new_empty(nq, k), I_parts.new_empty(nq, k)


def _flat_begin_cuda(q, handle: int, k: int):
    return _INDEXES[handle]._begin_impl(q, k)


def _flat_begin_fake(q, handle: int, k: int):
    return q.new_empty(q.shape[0], k)


def _flat_end_cuda(q, bound, handle: int, k: int, id_offset: int):
    return _INDEXES[handle]._end_impl(q, k, bound, id_offset)


def _flat_end_fake(q, bound, handle: int, k: int, id_offset: int):
    return q.new_empty(q.shape[0], k), q.new_empty(q.shape[0], k, dtype=torch.int64)


_define("flat_l2_begin(Tensor q, int index, int k) -> Tensor", _flat_begin_cuda, _flat_begin_fake)
_define("flat_l2_end(Tensor q, Tensor bound, int index, int k, int id_offset) -> (Tensor, Tensor)", _flat_end_cuda,
        _flat_end_fake)
_define("topk_merge(Tensor D_parts, Tensor I_parts) -> (Tensor, Tensor)", _merge_cuda, _merge_fake)


# ------------------------------------------------------------------------------------------------------------------
# building blocks of the launch plan (functional forms: they allocate their outputs)
# ------------------------------------------------------------------------------------------------------------------
def _conv_tc_cuda(src0, src1, w_hi, w_lo, bias, cout: int, act: int, dst_mode: int, dst_layout: int, pool: bool):
    """src0 / src1 channels-last (B,H,W,C); (w_hi, w_lo, bias) from ops.pack_conv_tc.  dst_mode 1: plain, 2:
    PixelShuffle(2); dst_layout 0: channels-last, 1: NCHW.  Returns (dst, pooled): pooled = MaxPool2d(2,2) of the
    result (channels-last) when ``pool``, else an empty tensor."""
    B, H, W, _ = src0.shape
    cpad = bias.numel()
    if dst_mode == 2:
        dst = src0.new_empty(B, 2 * H, 2 * W, cout // 4)
    elif dst_layout == 1:
        dst = src0.new_empty(B, cout, H, W)
    else:
        dst = src0.new_zeros(B, H, W, cpad) if cpad != cout else src0.new_empty(B, H, W, cpad)
    dpool = src0.new_empty(B, H // 2, W // 2, cpad) if pool else None
    hi = w_hi
    op = ops.tc_conv(src0, (hi, w_lo, bias), cout, act=act, src1=src1, dst=dst, dst_layout=dst_layout,
                     dst_mode=dst_mode, dst_pool=dpool)
    op.run()
    return dst, (dpool if pool else src0.new_empty(0))


def _conv_tc_fake(src0, src1, w_hi, w_lo, bias, cout: int, act: int, dst_mode: int, dst_layout: int, pool: bool):
    B, H, W, _ = src0.shape
    cpad = bias.numel()
    if dst_mode == 2:
        dst = src0.new_empty(B, 2 * H, 2 * W, cout // 4)
    elif dst_layout == 1:
        dst = src0.new_empty(B, cout, H, W)
    else:
        dst = src0.new_empty(B, H, W, cpad)
    return dst, (src0.new_empty(B, H // 2, W // 2, cpad) if pool else src0.new_empty(0))


_define("conv_tc(Tensor src0, Tensor? src1, Tensor w_hi, Tensor w_lo, Tensor bias, int cout, int act, int dst_mode, "
        "int dst_layout, bool pool) -> (Tensor, Tensor)", _conv_tc_cuda, _conv_tc_fake)


def _conv_rs_cuda(src0, src1, w_hi, w_lo, bias, w_scale: float, cout: int, act: int, dst_mode: int, dst_layout: int,
                  pool: bool):
    """One <= 64-channel slice of a 3x3 conv on the "3xFP16" row-stationary kernel (csrc/conv_rs.cu).  src0 / src1 and
    the channels-last outputs are in the SPLIT format (nanovs::split16 / unsplit16); (w_hi, w_lo, bias, w_scale) = one
    entry of ops.pack_conv_rs(...).slices.  dst_mode / dst_layout / pool as conv_tc."""
    B, H, W, _ = src0.shape
    cpad = bias.numel()
    if dst_mode == 2:
        dst = src0.new_zeros(B, 2 * H, 2 * W, (cout // 4 + 7) // 8 * 8)
    elif dst_layout == 1:
        dst = src0.new_empty(B, cout, H, W)
    else:
        dst = src0.new_zeros(B, H, W, cpad)
    dpool = src0.new_zeros(B, H // 2, W // 2, cpad) if pool else None
    op = ops.TcConv(src0, (w_hi, w_lo, bias), cout if dst_layout == 1 or dst_mode == 2 else cpad, act=act, src1=src1,
                    dst=dst, dst_layout=dst_layout, dst_mode=dst_mode, dst_pool=dpool, rs_scale=w_scale)
    op.run()
    return dst, (dpool if pool else src0.new_empty(0))


def _conv_rs_fake(src0, src1, w_hi, w_lo, bias, w_scale: float, cout: int, act: int, dst_mode: int, dst_layout: int,
                  pool: bool):
    B, H, W, _ = src0.shape
    cpad = bias.numel()
    if dst_mode == 2:
        dst = src0.new_empty(B, 2 * H, 2 * W, (cout // 4 + 7) // 8 * 8)
    elif dst_layout == 1:
        dst = src0.new_empty(B, cout, H, W)
    else:
        dst = src0.new_empty(B, H, W, cpad)
    return dst, (src0.new_empty(B, H // 2, W // 2, cpad) if pool else src0.new_empty(0))


_define("conv_rs(Tensor src0, Tensor? src1, Tensor w_hi, Tensor w_lo, Tensor bias, float w_scale, int cout, int act, "
        "int dst_mode, int dst_layout, bool pool) -> (Tensor, Tensor)", _conv_rs_cuda, _conv_rs_fake)
_define("split16(Tensor x) -> Tensor", lambda x: ops.split16(x), lambda x: torch.empty_like(x))
_define("unsplit16(Tensor x) -> Tensor", lambda x: ops.unsplit16(x), lambda x: torch.empty_like(x))


def _conv_cuda(src, weight, bias, cout: int, ksize: int, act: int):
    """NCHW conv (3x3 pad 1 or 1x1) on the exact-fp32 FFMA kernel; (weight, bias) from ops.pack_conv."""
    return ops.conv(src, weight, bias, cout, ksize=ksize, act=act)


def _conv_fake(src, weight, bias, cout: int, ksize: int, act: int):
    return src.new_empty(src.shape[0], cout, src.shape[2], src.shape[3])


_define("conv(Tensor src, Tensor weight, Tensor bias, int cout, int ksize, int act) -> Tensor", _conv_cuda, _conv_fake)

_define("attention(Tensor q, Tensor kv, int heads) -> Tensor",
        lambda q, kv, heads: ops.attention(q, kv, heads), lambda q, kv, heads: torch.empty_like(q))
_define("netvlad(Tensor x, Tensor w_assign, Tensor centroids) -> Tensor",
        lambda x, w, c: ops.netvlad(x, w, c), lambda x, w, c: x.new_empty(x.shape[0], c.shape[0] * x.shape[1]))
_define("channel_layernorm(Tensor x, Tensor g, Tensor b, float eps) -> Tensor",
        lambda x, g, b, eps: ops.channel_layernorm(x, g, b, eps), lambda x, g, b, eps: torch.empty_like(x))
_define("dwconv3x3(Tensor x, Tensor w, Tensor b) -> Tensor",
        lambda x, w, b: ops.dwconv3x3(x, w, b), lambda x, w, b: torch.empty_like(x))
_define("softmax_channels(Tensor x) -> Tensor",
        lambda x: ops.softmax_channels(x), lambda x: torch.empty_like(x))


def _prep_cuda(img, size: Sequence[int]):
    return ops.preprocess_u8(img, tuple(size) if size else None)


def _prep_fake(img, size: Sequence[int]):
    B, H, W, _ = img.shape
    Ho, Wo = (size[0], size[1]) if size else (H, W)
    return img.new_empty(B, 3, Ho, Wo, dtype=torch.float32)


_define("preprocess_u8(Tensor img, int[] size) -> Tensor", _prep_cuda, _prep_fake)
