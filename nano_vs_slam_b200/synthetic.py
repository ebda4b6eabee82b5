"""Deterministic synthetic weights and inputs (there are no checkpoints or datasets offline).

``spread_init`` is the fixture recipe from SURVEY.md §8(c): default PyTorch init gives scores in
[0, 0.51] (no keypoint passes 0.7) and near-tied segmentation logits, so parity fixtures use
He-normal convs and randomised BatchNorm affine/running statistics instead.  Values are generated
*per tensor name* (seed mixed with crc32 of the key), so any module tree that exposes the reference's
``state_dict`` keys and shapes receives bit-identical weights regardless of construction order.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


@torch.no_grad()
def spread_init(state_dict: Dict[str, torch.Tensor], seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Return a new fp32 CPU state_dict with spread-init values for every key of ``state_dict``."""
    out = {}
    for key in sorted(state_dict.keys()):
        ref = state_dict[key]
        shape = tuple(ref.shape)
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        parent = key.rsplit(".", 2)[-2] if key.count(".") >= 1 else ""
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.long)
        elif parent == "bn" and leaf == "weight":
            out[key] = torch.rand(shape, generator=g) + 0.5  # gamma ~ U(0.5, 1.5)
        elif parent == "bn" and leaf == "bias":
            out[key] = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_mean":
            out[key] = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            out[key] = torch.rand(shape, generator=g) + 0.5  # U(0.5, 1.5)
        elif leaf == "p" and shape == (1,):
            out[key] = 2.5 + torch.rand(shape, generator=g)  # GeM exponent (gem.py:11 default 3), kept positive
        elif leaf == "centroids":
            out[key] = torch.rand(shape, generator=g)  # netvlad.py:43 default
        elif parent == "norm" and leaf == "g":
            out[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif parent == "norm" and leaf == "b":
            out[key] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif leaf == "bias":
            out[key] = torch.randn(shape, generator=g) * 0.1
        else:
            out[key] = torch.randn(shape, generator=g) * 0.1
        if out[key].dtype != torch.long:
            out[key] = out[key].to(torch.float32)
    return out


def synthetic_frames(batch: int, H: int, W: int, seed: int = 0) -> torch.Tensor:
    """``rand(B,3,H,W)*2-1`` -- all reference loaders feed x*2-1 (frontend.py:79, pittsburgh.py:15)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.rand(batch, 3, H, W, generator=g) * 2.0 - 1.0


def planted_retrieval_set(n_db: int, n_q: int, dim: int, k: int, seed: int = 0,
                          device: str = "cpu", chunk: int = 65536):
    """Tie-free retrieval data (i.i.d. unit vectors are NOT tie-free in fp32).

    Queries are unit-normalised N(0, I).  For query i, ``k`` db rows j_0..j_{k-1} are *planted* as
    ``alpha_m * q_i + sqrt(1 - alpha_m^2) * r`` with r a unit vector orthogonal to q_i, so
    cos(q_i, db[j_m]) = alpha_m = 0.9 - 0.5*m/k exactly (up to rounding): adjacent squared-distance
    gaps are 1/k >= 1e-2 for k <= 100, while un-planted rows sit at d2 ~= 2 +- 0.2.
    Requires n_q * k <= n_db.  Returns (db, queries, planted) with planted[i, m] = j_m, the expected
    top-k index list of query i, in order.
    """
    assert n_q * k <= n_db, "need n_q*k <= n_db planted slots"
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    db = torch.empty(n_db, dim, device=device, dtype=torch.float32)
    for s0 in range(0, n_db, chunk):
        blk = torch.randn(min(chunk, n_db - s0), dim, generator=g, device=device)
        db[s0:s0 + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    q = torch.randn(n_q, dim, generator=g, device=device)
    q = q / q.norm(dim=1, keepdim=True)
    planted = torch.randperm(n_db, generator=g, device=device)[: n_q * k].view(n_q, k)
    alpha = (0.9 - 0.5 * torch.arange(k, device=device, dtype=torch.float32) / k).view(1, k, 1)
    qchunk = max(1, chunk // max(k, 1))
    for q0 in range(0, n_q, qchunk):
        qq = q[q0:q0 + qchunk]                                   # (b, dim)
        r = torch.randn(qq.shape[0], k, dim, generator=g, device=device)
        r = r - (r * qq.unsqueeze(1)).sum(-1, keepdim=True) * qq.unsqueeze(1)
        r = r / r.norm(dim=-1, keepdim=True)
        rows = alpha * qq.unsqueeze(1) + torch.sqrt(1 - alpha * alpha) * r
        rows = rows / rows.norm(dim=-1, keepdim=True)
        db[planted[q0:q0 + qchunk].reshape(-1)] = rows.reshape(-1, dim)
    return db, q, planted


def algorithmic_bytes_per_frame(model, H: int, W: int, with_decode: bool = True) -> float:
    """fp32 input + forward-dict outputs (+ decode outputs) of one frame -- the compulsory HBM traffic
    used as the roofline numerator; weights (<= 3.7 MB) amortise to 0 over a batch."""
    h2, w2 = H // 2, W // 2
    h4, w4 = h2 // 2, w2 // 2
    n = 3 * H * W + 3 * h4 * w4 + model.nfeatures * h2 * w2 + model.nClasses * h2 * w2 + model.get_global_desc_dim()
    if with_decode:
        n += (3 + model.nfeatures) * h4 * w4 + 2 * h2 * w2  # score, coord, sampled feat (f32) + seg argmax (i64)
    return 4.0 * n
