"""IndexFlatL2 -- exact L2 retrieval with faiss's call shape (src/evaluation/global_descriptor.py:55-60):

    index = IndexFlatL2(d); index.add(db); D, I = index.search(q, k)

``add`` keeps the fp32 rows resident in HBM and builds the fp16 GEMM operand, squared norms and the shard statistics
behind the screen's error bound once; ``search`` runs the tcgen05/TMEM GEMM with fused per-row lists, the selection
of everything inside the error bound, the fp32 re-rank and (for queries the lists cannot decide) the exact scan
(csrc/retrieval.cu): the result is the fp32 result.
numpy in -> numpy out (like faiss); CUDA tensors in -> CUDA tensors out.

ShardedIndexFlatL2 row-shards the database over the ranks of a torch.distributed process group: every rank runs the
GEMM on its shard, a small all_gather ((Q,k) floats per rank) agrees on a bound of the global k-th distance, every rank
re-ranks only its rows inside that bound (global ids), ONE all_gather moves the packed (Q,k) distances + labels (NCCL
over NVLink on GPUs), and every rank merges world_size*k candidates per query.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch

from . import ops, torch_ops
from ._cabi import check, lib

KMAX = 64  # nvs_flat_max_k(): per-row lists live in shared memory


def _as_dev(x, device) -> Tuple[torch.Tensor, bool]:
    was_np = isinstance(x, np.ndarray)
    if was_np:
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return x.to(device=device, dtype=torch.float32).contiguous(), was_np


class IndexFlatL2(object):
    def __init__(self, d: int, device="cuda"):
        self.d = int(d)
        self.device = torch.device(device)
        self.ntotal = 0
        self._x = None       # fp32 rows (re-rank operand, exact distances)
        self._xb = None      # fp16 rows padded to a multiple of 64 columns (GEMM operand), one scale per shard
        self._xn = None      # |x|^2
        self._st = None      # max |x|^2, max |x - fp16(x)|^2, scale exponent, max |component|
        self._ws = None

    def add(self, x) -> None:
        x, _ = _as_dev(x, self.device)
        assert x.dim() == 2 and x.shape[1] == self.d, (x.shape, self.d)
        self._x = x if self._x is None else torch.cat([self._x, x], 0)
        n = self._x.shape[0]
        dpad = int(lib().nvs_flat_padded_dim(self.d))
        self._xb = torch.empty(n, dpad, dtype=torch.float16, device=self.device)
        self._xn = torch.empty(n, dtype=torch.float32, device=self.device)
        self._st = torch.empty(4, dtype=torch.float32, device=self.device)
        check(lib().nvs_flat_prepare(self._x.data_ptr(), n, self.d, self._xb.data_ptr(), self._xn.data_ptr(),
                                     self._st.data_ptr(), ops._stream()), "nvs_flat_prepare")
        ops.LAUNCHES[0] += 2
        self.ntotal = n

    def search_device(self, q: torch.Tensor, k: int, id_offset: int = 0, gemm_events=None, out=None):
        """``gemm_events``: optional (start, stop) torch.cuda.Event pair recorded around the GEMM kernel.
        ``out``: optional (D (nq,k) float32, I (nq,k) int64) tensors that receive the result.
        Runs as the custom operator torch.ops.nanovs.flat_l2_search (CUDA dispatch key only)."""
        self._gemm_events, self._out = gemm_events, out
        try:
            return torch.ops.nanovs.flat_l2_search(q, torch_ops.register_index(self), int(k), int(id_offset))
        finally:
            self._gemm_events = self._out = None

    def _search_impl(self, q: torch.Tensor, k: int, id_offset: int = 0):
        gemm_events = getattr(self, "_gemm_events", None)
        assert self.ntotal > 0, "empty index"
        if k > KMAX:
            raise NotImplementedError(f"k <= {KMAX} (per-row lists live in shared memory)")
        if k > self.ntotal:
            raise ValueError(f"k = {k} > ntotal = {self.ntotal}")
        nq = q.shape[0]
        out = getattr(self, "_out", None)  # caller-provided (D, I) views, e.g. of one packed all_gather buffer
        if out is not None:
            D, I = out
        else:
            D = torch.empty(nq, k, dtype=torch.float32, device=self.device)
            I = torch.empty(nq, k, dtype=torch.int64, device=self.device)
        nbytes = int(lib().nvs_flat_search_workspace_bytes(self.ntotal, nq, self.d, k))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        e0 = e1 = None
        if gemm_events is not None:
            for ev in gemm_events:
                ev.record()  # forces creation of the underlying cudaEvent_t; re-recorded inside the library
            e0, e1 = gemm_events[0].cuda_event, gemm_events[1].cuda_event
        check(lib().nvs_flat_search(self._x.data_ptr(), self._xb.data_ptr(), self._xn.data_ptr(), self._st.data_ptr(),
                                    self.ntotal, q.data_ptr(), nq, self.d, k, id_offset, D.data_ptr(), I.data_ptr(),
                                    self._ws.data_ptr(), self._ws.numel(), e0, e1, ops._stream()),
              "nvs_flat_search")
        ops.LAUNCHES[0] += 6  # query conversion, bound fill, GEMM + lists, selection, re-rank, exact scan
        return D, I

    # --- two-phase search of one shard of a sharded database (nanovs.h: nvs_flat_search_begin / _end) -------------
    def _workspace(self, nq: int, k: int) -> torch.Tensor:
        """Workspace of slot ``self._slot`` (the lists travel in it from search_begin to search_end; a caller that
        pipelines query chunks over two streams alternates the slot)."""
        nbytes = int(lib().nvs_flat_search_workspace_bytes(self.ntotal, nq, self.d, k))
        slot = getattr(self, "_slot", 0)
        if slot == 0:
            if self._ws is None or self._ws.numel() < nbytes:
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            return self._ws
        extra = self.__dict__.setdefault("_ws_extra", {})
        if slot not in extra or extra[slot].numel() < nbytes:
            extra[slot] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return extra[slot]

    def _begin_impl(self, q: torch.Tensor, k: int) -> torch.Tensor:
        gemm_events = getattr(self, "_gemm_events", None)
        assert self.ntotal > 0, "empty index"
        if k > KMAX:
            raise NotImplementedError(f"k <= {KMAX} (per-row lists live in shared memory)")
        if k > self.ntotal:
            raise ValueError(f"k = {k} > ntotal = {self.ntotal}")
        nq = q.shape[0]
        ws = self._workspace(nq, k)
        bound = torch.empty(nq, k, dtype=torch.float32, device=self.device)
        e0 = e1 = None
        if gemm_events is not None:
            for ev in gemm_events:
                ev.record()
            e0, e1 = gemm_events[0].cuda_event, gemm_events[1].cuda_event
        check(lib().nvs_flat_search_begin(self._x.data_ptr(), self._xb.data_ptr(), self._xn.data_ptr(), self._st.data_ptr(),
                                          self.ntotal, q.data_ptr(), nq, self.d, k, bound.data_ptr(), ws.data_ptr(),
                                          ws.numel(), e0, e1, ops._stream()), "nvs_flat_search_begin")
        ops.LAUNCHES[0] += 4  # query conversion, bound fill, GEMM + lists, k-th value per query
        return bound

    def _end_impl(self, q: torch.Tensor, k: int, bound: torch.Tensor, id_offset: int = 0):
        nq = q.shape[0]
        out = getattr(self, "_out", None)
        if out is not None:
            D, I = out
        else:
            D = torch.empty(nq, k, dtype=torch.float32, device=self.device)
            I = torch.empty(nq, k, dtype=torch.int64, device=self.device)
        ws = self._workspace(nq, k)
        check(lib().nvs_flat_search_end(self._x.data_ptr(), self._xb.data_ptr(), self._xn.data_ptr(), self._st.data_ptr(),
                                        self.ntotal, q.data_ptr(), nq, self.d, k, id_offset, bound.data_ptr(),
                                        D.data_ptr(), I.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream()),
              "nvs_flat_search_end")
        ops.LAUNCHES[0] += 3  # selection, re-rank, exact scan
        return D, I

    def search_begin(self, q: torch.Tensor, k: int, gemm_events=None) -> torch.Tensor:
        """Phase 1 (torch.ops.nanovs.flat_l2_begin): GEMM + lists; returns (nq, k) upper bounds of the exact distances of
        this shard's k best rows -- gather them from all shards and reduce with merge_bounds before search_end."""
        self._gemm_events = gemm_events
        try:
            return torch.ops.nanovs.flat_l2_begin(q, torch_ops.register_index(self), int(k))
        finally:
            self._gemm_events = None

    def search_end(self, q: torch.Tensor, k: int, bound: torch.Tensor, id_offset: int = 0, out=None):
        """Phase 2 (torch.ops.nanovs.flat_l2_end): re-rank of the rows that can still be among the global k nearest."""
        self._out = out
        try:
            return torch.ops.nanovs.flat_l2_end(q, bound, torch_ops.register_index(self), int(k), int(id_offset))
        finally:
            self._out = None

    def search(self, q, k: int):
        qd, was_np = _as_dev(q, self.device)
        assert qd.dim() == 2 and qd.shape[1] == self.d
        D, I = self.search_device(qd, k)
        if was_np:
            return D.cpu().numpy(), I.cpu().numpy()
        return D, I


def merge_bounds(bounds: torch.Tensor) -> torch.Tensor:
    """(parts, nq, k) per-shard bounds of IndexFlatL2.search_begin -> (nq,) bound of the global k-th distance."""
    bounds = ops._req(bounds)
    parts, nq, k = bounds.shape
    out = torch.empty(nq, dtype=torch.float32, device=bounds.device)
    check(lib().nvs_flat_bound_merge(bounds.data_ptr(), parts, nq, k, out.data_ptr(), ops._stream()), "nvs_flat_bound_merge")
    ops.LAUNCHES[0] += 1
    return out


def merge_topk_device(D_parts: torch.Tensor, I_parts: torch.Tensor):
    """(parts, Q, k) sorted partial results -> (Q, k) global top-k (csrc/retrieval.cu merge_parts_kernel)."""
    return torch.ops.nanovs.topk_merge(D_parts, I_parts)


def _merge_topk_impl(D_parts: torch.Tensor, I_parts: torch.Tensor):
    parts, nq, k = D_parts.shape
    # any layout whose (nq, k) blocks are contiguous is read in place: e.g. views of one gathered buffer
    def _ok(t):
        return t.stride(2) == 1 and t.stride(1) == k and t.stride(0) >= nq * k
    if not _ok(D_parts):
        D_parts = D_parts.contiguous()
    if not _ok(I_parts):
        I_parts = I_parts.contiguous()
    D = torch.empty(nq, k, dtype=torch.float32, device=D_parts.device)
    I = torch.empty(nq, k, dtype=torch.int64, device=D_parts.device)
    check(lib().nvs_topk_merge(D_parts.data_ptr(), I_parts.data_ptr(), parts, nq, k, D_parts.stride(0),
                               I_parts.stride(0), D.data_ptr(), I.data_ptr(), ops._stream()), "nvs_topk_merge")
    ops.LAUNCHES[0] += 1
    return D, I


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of ``rank``: the first n_total % world ranks get one extra row."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedIndexFlatL2(object):
    """Row-sharded exact index: shard-local search + ONE all_gather + merge.

    ``local_search(shard, q, k, id_offset) -> (D, I)`` and ``merge(D_parts, I_parts) -> (D, I)`` default to the
    CUDA kernels; the CPU gloo tests inject the oracle's functions to exercise the host-side sharding logic.
    """

    def __init__(self, d: int, n_total: int, group=None, device="cuda",
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.d, self.n_total = d, n_total
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        self.device = torch.device(device)
        self._index = IndexFlatL2(d, device) if local_search is None else None
        self._local_search = local_search
        self._merge = merge if merge is not None else merge_topk_device
        self._shard = None

    def add_local(self, x_shard) -> None:
        """Rows [lo, hi) of the global database (each rank adds only its own shard)."""
        assert x_shard.shape[0] == self.hi - self.lo, (x_shard.shape, self.lo, self.hi)
        if self._index is not None:
            self._index.add(x_shard)
        else:
            self._shard = x_shard

    @staticmethod
    def _packed_views(buf: torch.Tensor, parts: int, nq: int, k: int):
        """(parts, stride) int64 words -> D (parts,nq,k) float32 and I (parts,nq,k) int64 views: a part is nq*k labels
        followed by nq*k distances (two per word)."""
        n = nq * k
        stride = buf.shape[1]
        I = buf.as_strided((parts, nq, k), (stride, k, 1), 0)
        D = buf.view(torch.float32).as_strided((parts, nq, k), (2 * stride, k, 1), 2 * n)
        return D, I

    def search(self, q: torch.Tensor, k: int, gemm_events=None):
        nq = q.shape[0]
        if self.world == 1:
            if self._index is not None:
                return self._index.search_device(q.to(self.device, torch.float32).contiguous(), k, id_offset=self.lo,
                                                 gemm_events=gemm_events)
            return self._local_search(self._shard, q, k, self.lo)
        if self._index is not None:
            return self._search_pipelined(q, k, gemm_events)
        # ONE packed buffer per rank (labels, then distances) -> ONE all_gather; the merge reads the gathered parts in place
        words = nq * k + (nq * k + 1) // 2
        dev = q.device
        mine = torch.empty(1, words, dtype=torch.int64, device=dev)
        Dm, Im = self._packed_views(mine, 1, nq, k)
        D, I = self._local_search(self._shard, q, k, self.lo)
        Dm[0].copy_(D)
        Im[0].copy_(I)
        gathered = torch.empty(self.world, words, dtype=torch.int64, device=dev)
        self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
        Dg, Ig = self._packed_views(gathered, self.world, nq, k)
        return self._merge(Dg, Ig)

    @staticmethod
    def query_chunks(nq: int, chunk: Optional[int] = None):
        """[a, b) query ranges of the pipelined sharded search: chunks of NVS_RETR_CHUNK queries, a last chunk below 512
        queries joins its predecessor.  Default: ONE chunk -- measured at 10k x 1M x 4096 (profiles/r2_retrieval_chunking.txt)
        the GEMM loses more on shorter launches (N = 8: 10.6 -> 13.1 ms, N = 4: 20.1 -> 22.3 ms) than the hidden
        exchange tail (1.1 - 2 ms) is worth."""
        import os

        chunk = int(os.environ.get("NVS_RETR_CHUNK", 1 << 30)) if chunk is None else chunk
        chunk = max(128, chunk)
        edges = list(range(0, nq, chunk)) + [nq]
        if len(edges) > 2 and edges[-1] - edges[-2] < 512:
            del edges[-2]
        return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]

    def _search_pipelined(self, q: torch.Tensor, k: int, gemm_events=None):
        """Sharded search on the CUDA kernels, pipelined over query chunks: the GEMM of chunk j + 1 (main stream) runs
        while chunk j's exchange tail -- all_gather of the per-shard k-th bounds, re-rank inside the GLOBAL bound,
        all_gather of the packed (labels, distances) parts, merge -- runs on a side stream.  Only the last chunk's tail
        is exposed.  ``gemm_events``: a list that receives one (start, stop) event pair per chunk, or one pair
        (recorded around the first chunk only)."""
        dev = self.device
        qd = q.to(dev, torch.float32).contiguous()
        nq = qd.shape[0]
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=dev)
        side = self._side
        chunks = self.query_chunks(nq)
        keep, outs, done = [], [], []
        for j, (a, b) in enumerate(chunks):
            n = b - a
            qj = qd[a:b]
            ge = None
            if isinstance(gemm_events, list):
                ge = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                gemm_events.append(ge)
            elif gemm_events is not None and j == 0:
                ge = gemm_events
            self._index._slot = j % 2
            if j >= 2:
                main.wait_event(done[j - 2])  # the workspace slot is free again
            mine_b = self._index.search_begin(qj, k, gemm_events=ge)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                all_b = torch.empty(self.world, n, k, dtype=torch.float32, device=dev)
                self.dist.all_gather_into_tensor(all_b, mine_b, group=self.group)
                words = n * k + (n * k + 1) // 2
                mine = torch.empty(1, words, dtype=torch.int64, device=dev)
                Dm, Im = self._packed_views(mine, 1, n, k)
                self._index._slot = j % 2
                self._index.search_end(qj, k, merge_bounds(all_b), id_offset=self.lo, out=(Dm[0], Im[0]))
                gathered = torch.empty(self.world, words, dtype=torch.int64, device=dev)
                self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
                Dg, Ig = self._packed_views(gathered, self.world, n, k)
                outs.append(self._merge(Dg, Ig))
                d = torch.cuda.Event()
                d.record(side)
                done.append(d)
            keep.append((qj, mine_b, all_b, mine, gathered))  # alive until the main stream has waited for the side stream
        self._index._slot = 0
        main.wait_stream(side)
        for D, I in outs:
            D.record_stream(main)
            I.record_stream(main)
        if len(outs) == 1:
            return outs[0]
        D = torch.cat([o[0] for o in outs], 0)
        I = torch.cat([o[1] for o in outs], 0)
        del keep
        return D, I
