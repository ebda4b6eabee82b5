"""KP2DtinyFrontend -- device-side version of src/visual_odometry/frontend.py:11-129.

The reference moves score / coord / feat / seg to the host every frame and does the threshold,
optional semantic filter and argpartition top-k in numpy (frontend.py:94-126).  Here the whole decode
(post_processing + selection) stays on the GPU; ``run`` keeps the reference's single-image signature and
numpy return types, ``run_batch`` is the batched form used for throughput (frames are independent, the
reference glue is simply applied per frame).
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Iterable, Iterator, Optional, Sequence

import numpy as np
import torch

from . import ops, torch_ops
from .kp2dtiny import tiny_factory


class KP2DtinyFrontend(object):
    def __init__(self, new_size=None, weights_path=None, nn_thresh=0.7, device="cuda", semantic_filter=False,
                 classes_to_filter=(21,), debug=False, method="kp2dtiny", config="S", top_k=4000, v3=False,
                 nClasses=28, state_dict=None):
        if method != "kp2dtiny":
            raise NotImplementedError("only the kp2dtiny extractor is on the B200 hot path")
        self.name = method
        self.device = device
        self.nn_thresh = nn_thresh
        self.border_remove = 4
        self.classes_to_filter = list(classes_to_filter)
        self.apply_semantic_filer = semantic_filter  # (sic) attribute name of the reference, frontend.py:34
        self.weights_path = weights_path
        self.plot = False  # cv2.imshow debugging of the reference (:95-105) is not part of the hot path
        self.top_k = top_k
        self.new_size = new_size
        self.net = tiny_factory(config, nClasses, to_export=False, to_mcu=False, v3=v3)
        self.net.sample_segmentation = semantic_filter  # frontend.py:47
        if weights_path is not None:
            state_dict = torch.load(weights_path, map_location=torch.device("cpu"))["state_dict"]  # :51-53
        if state_dict is not None:
            self.net.load_state_dict(state_dict)
        self.net.eval()
        self.net.training = False
        self.net = self.net.to(self.device)
        self.net.device = self.device

    def get_info(self):
        return {"nn_thresh": self.nn_thresh, "border_remove": self.border_remove, "weights_path": self.weights_path,
                "device": self.device, "apply_semantic_filer": self.apply_semantic_filer,
                "classes_to_filter": self.classes_to_filter, "plot": self.plot, "top_k": self.top_k,
                "new_size": self.new_size, "name": self.name, "model": self.net.gather_info()}

    @torch.no_grad()
    def run_batch(self, imgs: torch.Tensor, normalized: bool = False):
        """imgs (B,3,H,W) in [0,1] (or already in [-1,1] if ``normalized``), or uint8 (B,H,W,3) camera frames,
        on any device.

        Returns (sel, post): ``sel`` = dict(pts, desc, score, cell, label, count) device tensors from
        ops.select_keypoints, ``post`` = the post_processing dict (vlad, seg argmax, dense maps).
        """
        x = imgs.to(self.device, non_blocking=True)
        if x.dtype == torch.uint8:
            # camera frames (B,H,W,3) uint8: /255, resize to new_size and (x-0.5)*2 happen on the device, fused
            # into the first conv's load stage when no resize is needed (visual_odometry.py:281-291)
            B, Hi, Wi, _ = x.shape
            if self.new_size is not None and tuple(self.new_size) != (Hi, Wi):
                x = torch.ops.nanovs.preprocess_u8(x, list(self.new_size))
                H, W = x.shape[2:]
            else:
                H, W = Hi, Wi
            unit = False
        else:
            unit = not normalized  # [0,1] frames: x.sub(0.5).mul(2.0) (frontend.py:79) happens in the stem kernel's load
            _, _, H, W = x.shape
        out = self.net.forward(x, unit_input=unit)
        post = self.net.post_processing(out, H, W)
        seg_cells = post["seg"] if self.apply_semantic_filer else None
        sel = torch_ops.select_keypoints(post["score"], post["coord"], post["feat"], self.nn_thresh, self.top_k,
                                         seg_cells=seg_cells,
                                         classes_to_filter=self.classes_to_filter if self.apply_semantic_filer else None)
        return sel, post

    @torch.no_grad()
    def stream(self, host_batches: Iterable[torch.Tensor], normalized: bool = True,
               with_seg: bool = True) -> Iterator[Dict[str, torch.Tensor]]:
        """Host-to-host streaming: pinned (B,3,H,W) batches in, dicts of pinned host tensors out.

        The reference front-end does `.to(device)` -> forward -> `.cpu()` serially per frame
        (frontend.py:81-116).  Here the same three phases are software-pipelined over three CUDA streams
        (H2D of batch i+1 and D2H of batch i-1 overlap the kernels of batch i; B200 has independent copy
        engines per direction), so the host always gets every batch's results, one batch behind the GPU.
        Yields, per batch: pts (B,k,2), desc (B,k,D), score (B,k), count (B,), vlad (B,G) [, seg (B,1,H/2,W/2)].
        The yielded tensors are views of three pinned result sets owned by this object and used in rotation: the
        device-to-host copy of batch s+3 reuses the set of batch s and is enqueued when the consumer asks for the
        batch after s+1, so a result may be held across ONE further next() (a = next(g); b = next(g); use(a) is
        safe), not longer (also across stream() calls) -- copy what must live longer."""
        dev = torch.device(self.device)
        comp = torch.cuda.current_stream(dev)
        h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # three pinned result sets, rotated and kept across stream() calls (cudaHostAlloc of ~80 MB costs tens of
        # milliseconds): one being filled, one just yielded, one the consumer may still hold from the yield before
        host_sets: list = self.__dict__.setdefault("_host_sets", [])
        # three device input buffers, rotated and kept across calls: nothing on this path goes through
        # record_stream (which defers block reuse and makes the caching allocator fall back to cudaMalloc /
        # cudaFree -- a device-wide stall -- every now and then)
        dev_in: list = self.__dict__.setdefault("_dev_in", [])
        slot_free: list = [None, None, None]  # compute-stream event after which input slot i may be overwritten

        def upload(hb, slot):
            if len(dev_in) <= slot or dev_in[slot].shape != hb.shape or dev_in[slot].dtype != hb.dtype:
                buf = torch.empty(hb.shape, dtype=hb.dtype, device=dev)
                if len(dev_in) <= slot:
                    dev_in.append(buf)
                else:
                    dev_in[slot] = buf
                # the block comes from the compute stream's pool and may have been freed by tensors whose kernels
                # are still queued there: the first copy into it must wait for them (once per slot and shape)
                h2d.wait_stream(comp)
            with torch.cuda.stream(h2d):
                if slot_free[slot] is not None:
                    h2d.wait_event(slot_free[slot])
                dev_in[slot].copy_(hb, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d)
            return dev_in[slot], ev

        it = iter(host_batches)
        try:
            nxt = upload(next(it), 0)
        except StopIteration:
            return
        try:
            yield from self._stream_loop(it, nxt, upload, slot_free, host_sets, comp, d2h, normalized, with_seg)
        finally:
            # a consumer that abandons the generator leaves a prefetch in flight: the persistent input buffers
            # must be quiescent before the next call (or the allocator) touches them
            h2d.synchronize()
            d2h.synchronize()

    def _stream_loop(self, it, nxt, upload, slot_free, host_sets, comp, d2h, normalized, with_seg):
        pending: deque = deque()
        step = 0
        while nxt is not None:
            x, ev = nxt
            try:
                nxt = upload(next(it), (step + 1) % 3)  # prefetch the next batch while this one computes
            except StopIteration:
                nxt = None
            comp.wait_event(ev)
            sel, post = self.run_batch(x, normalized=normalized)
            done = torch.cuda.Event()
            done.record(comp)
            slot_free[step % 3] = done
            dev_out = {"pts": sel["pts"], "desc": sel["desc"], "score": sel["score"], "count": sel["count"],
                       "vlad": post["vlad"]}
            if with_seg:
                dev_out["seg"] = post["seg"]
            slot = self.__dict__.get("_host_rr", 0)  # rotation continues across stream() calls
            self._host_rr = (slot + 1) % 3
            if len(host_sets) <= slot or any(k not in host_sets[slot] or host_sets[slot][k].shape != v.shape or
                                             host_sets[slot][k].dtype != v.dtype for k, v in dev_out.items()):
                fresh = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in dev_out.items()}
                if len(host_sets) <= slot:
                    host_sets.append(fresh)
                else:
                    host_sets[slot] = fresh
            host = host_sets[slot]
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                for k, v in dev_out.items():
                    host[k].copy_(v, non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(d2h)
            pending.append((host, fin, dev_out))  # dev_out stays referenced until its copy has finished
            step += 1
            if len(pending) > 1:
                h, e, _keep = pending.popleft()
                e.synchronize()
                yield h
        while pending:
            h, e, _keep = pending.popleft()
            e.synchronize()
            yield h

    def run(self, img):
        """Reference signature (frontend.py:78-129): img (3,H,W) in [0,1] -> (pts (n,2), desc (n,D), seg).
        Also accepts the raw camera frame, uint8 (H,W,3) numpy array or tensor (what VisualOdometry.process_image
        receives, visual_odometry.py:281): conversion, resize to ``new_size`` and normalisation then run on the GPU."""
        if isinstance(img, np.ndarray):
            img = torch.from_numpy(np.ascontiguousarray(img))
        sel, post = self.run_batch(img.unsqueeze(0))
        n = int(sel["count"][0])
        pts = sel["pts"][0, :n].cpu().numpy()
        feat = sel["desc"][0, :n].cpu().numpy()
        if self.apply_semantic_filer:
            seg = sel["label"][0, :n].cpu().numpy()
        else:
            # the reference indexes the H/2 x W/2 label map by cell index when more than top_k points pass
            # (frontend.py:116,126: "harmless quirk"); mirrored for drop-in equality
            seg = post["seg"].view(-1).cpu().numpy()
            if int((post["score"] > self.nn_thresh).sum()) > self.top_k > 0:
                seg = seg[sel["cell"][0, :n].cpu().numpy()]
        return pts.copy(), feat.copy(), seg.copy()
