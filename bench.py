#!/usr/bin/env python
"""bench.py -- KP2DTiny-S frames/s @240x320 on N B200 (+ NetVLAD retrieval queries/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the perception hot path over one batch of synthetic frames that is already
resident in HBM: KP2DTinyV2("S").forward -> post_processing -> keypoint selection (thr 0.7, top-1000).
Frames are independent, so for N > 1 every rank processes its own batch (weak scaling, no data-path
collective); the timed region is bracketed by barrier + synchronize and the max over ranks is used.
One JSON line is printed by rank 0 (contract: see the task statement / DESIGN.md §5).

--impl reference times the CPU restatement of the reference path (oracle/, torch CPU kernels == the
reference's own backend) on the host cores with the same metric/config; rank 0 only.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# ~1.5 GB of fresh output tensors per step go through the caching allocator; with fixed-size segments a new stream
# occasionally needs a cudaMalloc of that size in the middle of the timed region (a 40 ms stall seen in the per-batch
# trace).  Expandable segments grow in place instead.  (Must be set before CUDA is initialised.)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
# stdout carries the ONE JSON line and nothing else.  NCCL writes its NCCL_DEBUG=INFO/VERSION lines to the process's
# stdout from native code; they are NOT silenced (the driver counts ranks from them): file descriptor 1 is pointed at
# stderr for the whole run and the JSON line goes to a private duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(line: dict) -> None:
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


H, W = 240, 320
LETTER, V3, NCLS = "S", False, 28
THRESH, TOPK = 0.7, 1000
WSEED, XSEED = 1234, 0
FP32_FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # nominal, at max SM clock


def _peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


def _tf32_peak(bf16_peak: float):
    """Dense TF32 tensor peak: measured on this pool with cuBLAS (tools/measure_tf32_peak.py ->
    profiles/r2_tf32_peak.json, the sustained figure: kernels here are timed inside a long step), else bf16 / 2."""
    path = os.path.join(REPO, "profiles", "r2_tf32_peak.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d["tf32_tflops_sustained"], "measured cuBLAS tf32 sustained (profiles/r2_tf32_peak.json)"
    return bf16_peak / 2.0, "bf16 sustained / 2 (tf32 not measured)"


def _conv_tensor_peak(shape: str, bf16_peak: float):
    """Ceiling for ALGORITHMIC conv FLOP/s of a tensor-core conv launch: every algorithmic FLOP costs three tensor-core
    FLOPs (hi x hi + the two correction products).  "rs/f16" launches (csrc/conv_rs.cu) run kind::f16 MMAs: measured
    cuBLAS bf16 sustained / 3; the 3xTF32 kernels (csrc/conv_tc.cu): measured cuBLAS tf32 sustained / 3."""
    if "rs/f16" in shape:
        return bf16_peak / 3.0, "measured cuBLAS bf16 sustained (MEASURED_PEAKS.json; fp16 MMAs run at the bf16 rate) / 3 (3xFP16 split)", "f16"
    tf32_peak, tf32_src = _tf32_peak(bf16_peak)
    return tf32_peak / 3.0, tf32_src + " / 3 (3xTF32 split)", "tf32"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_model(device):
    from nano_vs_slam_b200 import tiny_factory
    from nano_vs_slam_b200.synthetic import spread_init

    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(LETTER, NCLS, v3=V3)
    sd = spread_init(m.state_dict(), WSEED)
    m.load_state_dict(sd)
    m.eval()
    m.training = False
    return m.to(device), sd


def workload_config(batch: int, world: int, backend: str) -> dict:
    """The `config` object of the JSON line -- identical for the native and the reference arm."""
    return {"workload": f"KP2DTiny-{LETTER} ({'V3 decoder fusion' if V3 else 'V2 dedicated decoders'}, {NCLS} classes) forward + post_processing + "
                        f"keypoint select (thr {THRESH}, top-{TOPK}), batch {batch} x {H}x{W} per GPU",
            "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"frame-dp{world}",
            "conv_backend": backend,
            "l2": "two alternating resident input batches of 236 MB each (> 126 MB L2); activations "
                  "per step ~10 GB"}


def reference_step(sd, sample_frames: int):
    """-> (step(), kind, what): one pass of the reference's CPU path over ``sample_frames`` frames.

    kind "reference": the reference's OWN modules (oracle/_ref, staged unmodified by oracle/make_ref.py) run
    ``model(x)`` + ``model.post_processing`` as its callers do (eval_multitask.py:195-198, frontend.py:84-85); the
    threshold / top-k glue of frontend.py:94-126 is its literal numpy restatement (frontend.py itself imports modules
    that are not installed here).  kind "port": the functional restatement in oracle/ when oracle/_ref is absent."""
    from oracle import glue_ref, make_ref
    from nano_vs_slam_b200.synthetic import synthetic_frames

    x = synthetic_frames(sample_frames, H, W, XSEED)
    nfeat = 32

    def glue(post):
        for b in range(sample_frames):
            one = {k: post[k][b:b + 1] for k in ("score", "coord", "feat", "seg")}
            glue_ref.frontend_decode(one, one["feat"].shape[1], THRESH, TOPK)

    ref = None
    try:
        ref = make_ref.import_reference()
    except Exception as exc:  # a broken staging must not take the bench down: say so and use the port
        print(f"[bench] oracle/_ref not importable ({exc!r}); timing the oracle port instead", file=sys.stderr)
    if ref is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref.tiny_factory(LETTER, NCLS, v3=V3)
        m.load_state_dict(sd)
        m.eval()
        m.training = False

        def step():
            with torch.no_grad():
                out = m(x)
                glue(m.post_processing(out, H, W))

        return step, "reference", "the reference's own nn.Module (oracle/_ref, unmodified) on torch CPU (oneDNN)"
    from oracle import kp2dtiny_ref as R

    a = R.arch_for(LETTER, V3, NCLS)

    def step():
        out = R.forward(x, sd, a)
        glue(R.post_processing(out, H, W, a))

    return step, "port", "torch CPU (oneDNN) restatement in oracle/ (oracle/_ref not staged)"


def cpu_reference_fps(sd, sample_frames: int, iters: int, threads: int, min_seconds: float = 0.0):
    """The reference's CPU path on a bounded sample: ``iters`` steps, continuing until ``min_seconds`` of CPU work
    have been timed.  -> (frames/s, seconds, kind, what)."""
    torch.set_num_threads(threads)
    step, kind, what = reference_step(sd, sample_frames)
    step()
    t0 = time.perf_counter()
    done = 0
    while done < iters or (time.perf_counter() - t0) < min_seconds:
        step()
        done += 1
    dt = time.perf_counter() - t0
    return sample_frames * done / dt, dt, kind, what


def _traffic(key: str):
    """ncu --set full DRAM bytes per launch recorded under profiles/ (newest round first), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(REPO, "profiles", name)
        if os.path.exists(path):
            with open(path) as fh:
                d = json.load(fh)
            if key in d:
                return d[key]
    return None


def retrieval_cpu_baseline(threads: int, n_q: int = 256, n_db: int = 100_000, dim: int = 4096, k: int = 25,
                           min_seconds: float = 8.0):
    """the fp32 restatement of IndexFlatL2 (oracle/glue_ref.flat_l2_search; faiss is not installable
    here) on a 256-query x 100k-row x 4096-d slice, scaled linearly in the row count to the 1M-row database."""
    from oracle import glue_ref

    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    db = torch.randn(n_db, dim, generator=g)
    q = torch.randn(n_q, dim, generator=g)
    glue_ref.flat_l2_search(db, q[:32], k)
    t0 = time.perf_counter()
    done = 0
    while done < 1 or time.perf_counter() - t0 < min_seconds:
        glue_ref.flat_l2_search(db, q, k)
        done += 1
    dt = (time.perf_counter() - t0) / done
    full = int(os.environ.get("NVS_RETR_NDB", 1_000_000))
    return {"value": n_q / (dt * full / n_db), "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{n_q} queries x {n_db} rows x {dim}-d fp32, top-{k}, {done} searches of {dt:.2f} s each on the "
                      f"host cores, scaled linearly to {full} rows (torch CPU restatement of IndexFlatL2 in oracle/)"}


def run_other_configs(dev, world, rank, hbm_peak, bf16_peak, which):
    """BASELINE.json configs 2, 3 and 4 on the same kernels, device-timed like the headline (>= 3 warm-up steps,
    barrier + synchronize on both sides, max over ranks, two alternating resident input batches larger than L2)."""
    import torch.distributed as dist
    from nano_vs_slam_b200 import tiny_factory, torch_ops
    from nano_vs_slam_b200.matcher import pose_consecutive
    from nano_vs_slam_b200.synthetic import algorithmic_bytes_per_frame, spread_init, synthetic_frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kitti_k = (718.856, 718.856, 607.19, 185.22)  # KITTI sequence 00 pinhole intrinsics
    specs = [
        dict(config=2, letter="N", v3=True, ncls=28, h=240, w=320, batch=256, steps=5, thresh=0.7, topk=1000),
        # config 3: under spread-init weights no V2-S score passes 0.7, so the front-end runs top-k only
        dict(config=3, letter="S", v3=False, ncls=19, h=376, w=1241, batch=32, steps=5, thresh=0.0, topk=4000, vo=True),
        dict(config=4, letter="S_A", v3=False, ncls=19, h=512, w=1024, batch=64, steps=3, thresh=0.7, topk=1000),
    ]
    out = []
    for sp in specs:
        with contextlib.redirect_stdout(io.StringIO()):
            m = tiny_factory(sp["letter"], sp["ncls"], v3=sp["v3"])
        m.load_state_dict(spread_init(m.state_dict(), WSEED))
        m.eval()
        m.training = False
        m = m.to(dev)
        h, w, b = sp["h"], sp["w"], sp["batch"]
        if sp.get("vo"):
            # consecutive frames: frame t+1 = frame t shifted by two pixels + noise, so that matches exist
            xs = []
            for i in range(2):
                x0 = synthetic_frames(1, h, w + 2 * b, XSEED + 100 * rank + i)
                g = torch.Generator().manual_seed(i)
                xs.append(torch.cat([x0[:, :, :, 2 * j:2 * j + w] + 0.01 * torch.randn(1, 3, h, w, generator=g)
                                     for j in range(b)]).to(dev))
        else:
            xs = [synthetic_frames(b, h, w, XSEED + 100 * rank + i).to(dev) for i in range(2)]

        def step(x):
            o = m(x)
            post = m.post_processing(o, h, w)
            sel = torch_ops.select_keypoints(post["score"], post["coord"], post["feat"], sp["thresh"], sp["topk"])
            if sp.get("vo"):  # ratio-0.7 one-to-one matching + relative pose of the consecutive pairs, on the device
                return pose_consecutive(sel, kitti_k)
            return sel

        for i in range(3):
            step(xs[i % 2])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(sp["steps"]):
            step(xs[i % 2])
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        plan = next(iter(m._plans.values()))
        flops = sum(v["flops"] for v in plan.meta.values()) / b
        if m.use_attention:  # 4 Nq Nk C per attention block: H/4 and H/8 maps, keys at half resolution
            c5 = m.channel_dims[4]
            n1, n2 = (h // 4) * (w // 4), (h // 8) * (w // 8)
            flops += 4.0 * c5 * (n1 * (n1 // 4) + n2 * (n2 // 4))
        fps = world * b * sp["steps"] / (ms / 1e3)
        fps_gpu = fps / world
        tc_shapes = [v["shape"] for v in plan.meta.values() if "tcgen05" in v.get("shape", "")]
        peak, peak_src, _ = _conv_tensor_peak(tc_shapes[0] if tc_shapes else "", bf16_peak)
        out.append({
            "config": sp["config"], "value": fps, "unit": "frames/s", "ms_per_step": ms / sp["steps"],
            "steps": sp["steps"], "warmup": 3, "batch_per_gpu": b,
            "workload": f"KP2DTiny-{sp['letter']} ({'V3' if sp['v3'] else 'V2'}, {sp['ncls']} classes) forward + "
                        f"post_processing + keypoint select (thr {sp['thresh']}, top-{sp['topk']})"
                        + (" + ratio-0.7 one-to-one matching + five-point relative pose of the consecutive pairs"
                           if sp.get("vo") else "") + f", batch {b} x {h}x{w} per GPU",
            "roofline": {"bound": "tensor", "achieved": fps_gpu * flops / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": fps_gpu * flops / 1e12 / peak, "traffic": None,
                         "peak_source": peak_src,
                         "note": "whole step: algorithmic conv" + (" + attention" if m.use_attention else "") +
                                 " FLOPs per frame x frames/s (the attention core and the first layer run on the fp32 "
                                 "pipe, so the tensor ceiling is an upper bound for them)",
                         "hbm_frac_at_algorithmic_bytes":
                             fps_gpu * algorithmic_bytes_per_frame(m, h, w, True) / (hbm_peak * 1e9)},
        })
        del m, xs, plan
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nano_vs_slam_b200.synthetic import spread_init
    from nano_vs_slam_b200 import tiny_factory

    with contextlib.redirect_stdout(io.StringIO()):
        m = tiny_factory(LETTER, NCLS, v3=V3)
    sd = spread_init(m.state_dict(), WSEED)
    threads = os.cpu_count() or 1
    sample = 8  # frames per step: bounded sample of the batch-256 workload
    torch.set_num_threads(threads)
    step, kind, what = reference_step(sd, sample)
    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": f"KP2DTiny-{LETTER} frames/s @{H}x{W}", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the native arm's config, verbatim; this arm times a bounded sample of that workload on the host cores
        # (frames are independent: frames/s does not depend on the batch), described under cpu_baseline
        "config": workload_config(args.batch, max(1, args.gpus), os.environ.get("NVS_CONV_BACKEND", "tc")),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} frames per step x {args.steps} steps of the batch-{args.batch} workload, "
                                   f"{what}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    global LETTER, V3, NCLS, H, W
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--letter", default=LETTER, help="config letter (S, S_A, N, N_A); headline metric is S")
    ap.add_argument("--v3", action="store_true", help="KP2DTinyV3 (decoder fusion) instead of V2")
    ap.add_argument("--classes", type=int, default=NCLS)
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE config 2-4 sub-records")
    args = ap.parse_args()
    LETTER, V3, NCLS, H, W = args.letter, args.v3, args.classes, args.height, args.width
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from nano_vs_slam_b200 import ops
    from nano_vs_slam_b200.synthetic import synthetic_frames

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(3, args.warmup)
    B = args.batch

    model, sd = build_model(dev)
    # two distinct resident input batches (2 x 236 MB at B=256: larger than the 126 MB L2), alternated
    xs = [synthetic_frames(B, H, W, XSEED + 100 * rank + i).to(dev) for i in range(2)]

    def step(x):
        out = model(x)
        post = model.post_processing(out, H, W)
        sel = ops.select_keypoints(post["score"], post["coord"], post["feat"], THRESH, TOPK)
        return sel, post

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi needs ~0.5 s to start: sample from warm-up on
    for i in range(warm):
        step(xs[i % 2])
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 1.0:  # keep the GPU under load while the sampler spins up (untimed)
        step(xs[0])
        torch.cuda.synchronize()
    plan = next(iter(model._plans.values()))
    heavy = max(plan.meta, key=lambda i: plan.meta[i]["flops"])
    barrier()
    l0 = ops.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(xs[i % 2])
    e1.record()
    barrier()
    launches = ops.LAUNCHES[0] - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # dominant kernel: CUDA events around that one launch, live, over a second pass of the same steps (kept out
    # of the timed region so that the event records cannot perturb `value`; small batches replay a CUDA graph in
    # the timed region and run eagerly here)
    plan.profile = {"idx": heavy, "events": []}
    for i in range(max(3, min(args.steps, 10))):
        step(xs[i % 2])
    torch.cuda.synchronize()
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in plan.profile["events"])
    heavy_meta = plan.meta[heavy]
    plan.profile = None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers ("e2e") ----
    # KP2DtinyFrontend.stream: pinned host frames in -> pinned host keypoints/descriptors/VLAD/labels out, every
    # step; H2D / kernels / D2H of consecutive batches overlap on three CUDA streams (copies are inside the
    # timed region, the host consumes every batch's results).
    from nano_vs_slam_b200.frontend import KP2DtinyFrontend
    fe = KP2DtinyFrontend(config=LETTER, v3=V3, nClasses=NCLS, nn_thresh=THRESH, top_k=TOPK, device=dev,
                          state_dict=sd)
    host_x = [synthetic_frames(B, H, W, XSEED + 100 * rank + i).pin_memory() for i in range(2)]
    h2d = host_x[0].numel() * 4

    trace = os.environ.get("NVS_BENCH_TRACE") == "1"  # per-batch host arrival times of the e2e stream, to stderr

    def run_stream(n, arrivals=None):
        d2h_bytes, got = 0, 0
        t_prev = time.perf_counter()
        for res in fe.stream((host_x[i % 2] for i in range(n)), normalized=True):
            t_now = time.perf_counter()
            if arrivals is not None:
                arrivals.append(t_now)
            if trace:
                print(f"[e2e] batch {got}: +{1e3 * (t_now - t_prev):.1f} ms", file=sys.stderr)
            t_prev = t_now
            got += int(res["count"][0] >= 0)  # the host reads the results of every batch
            d2h_bytes = sum(t.numel() * t.element_size() for t in res.values())
        assert got == n
        return d2h_bytes

    import gc
    run_stream(3)
    gc.collect()
    gc.disable()  # no collector pauses inside the host-driven timed regions below
    # three back-to-back passes of K batches, each bracketed by barrier + synchronize and timed on the device (max over
    # ranks); `value` is frames / time over ALL passes, and the per-batch host arrival intervals (median / max, rank 0)
    # tell a uniformly slow stream from one stall (a 20-batch mean cannot)
    E2E_PASSES = 3
    pass_ms, intervals = [], []
    for _ in range(E2E_PASSES):
        barrier()
        arrivals = []
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        d2h = run_stream(args.steps, arrivals)
        f1.record()
        barrier()
        t_pass = f0.elapsed_time(f1)  # device clock around the whole host-driven stream (results all on the host)
        if world > 1:
            t = torch.tensor([t_pass], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_pass = float(t)
        pass_ms.append(t_pass)
        intervals += [1e3 * (b - a) for a, b in zip(arrivals[:-1], arrivals[1:])]
    e2e_ms = sum(pass_ms) / E2E_PASSES
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    e2e_stats = {"passes": E2E_PASSES, "frames_per_s_per_pass": [world * B * args.steps / (t / 1e3) for t in pass_ms],
                 "batch_interval_ms_median": statistics.median(intervals) if intervals else None,
                 "batch_interval_ms_max": max(intervals) if intervals else None}

    # the same stream fed with uint8 HWC camera frames (what the VO loop gets from the decoder,
    # visual_odometry.py:281): 1/4 of the H2D bytes, /255 and (x-0.5)*2 fused into the stem kernel's load
    host_u8 = [((hx.permute(0, 2, 3, 1) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
               for hx in host_x]

    def run_stream_u8(n):
        got = 0
        for res in fe.stream((host_u8[i % 2] for i in range(n))):
            got += int(res["count"][0] >= 0)
        assert got == n

    run_stream_u8(3)
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    run_stream_u8(args.steps)
    u1.record()
    barrier()
    u8_ms = u0.elapsed_time(u1)
    if world > 1:
        t = torch.tensor([u8_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u8_ms = float(t)
    gc.enable()
    e2e_u8 = {"value": world * B * args.steps / (u8_ms / 1e3), "unit": "frames/s",
              "h2d_bytes_per_step": host_u8[0].numel(), "note": "uint8 HWC frames in, same outputs"}

    # ---- roofline of the dominant kernel + whole-step figures ----
    from nano_vs_slam_b200.synthetic import algorithmic_bytes_per_frame
    flops_frame = sum(m["flops"] for m in plan.meta.values()) / B  # conv FLOPs (97 % of the model)
    bytes_frame = algorithmic_bytes_per_frame(model, H, W, with_decode=True)
    hbm_peak, bf16_peak, which = _peaks()
    fps_gpu = value / world
    ach_gbs = heavy_meta["bytes"] / (kern_ms / 1e3) / 1e9
    ach_tf = heavy_meta["flops"] / (kern_ms / 1e3) / 1e12
    traffic, traffic_batch = None, None
    tr = _traffic("conv_rs_96_64_120x160" if "rs/f16" in heavy_meta["shape"] else "conv_tc_96_64_120x160")
    if tr and "96->64 k3 @120x160" in heavy_meta["shape"]:
        traffic = tr["dram_bytes_per_launch"] * B / tr["batch"]  # ncu capture at batch 256, linear in batch
        traffic_batch = tr["batch"]
    if "tcgen05" in heavy_meta["shape"]:
        # three tensor-core FLOPs per algorithmic FLOP: the ceiling for ALGORITHMIC FLOP/s is the measured dense peak
        # of the MMA kind / 3
        peak_tf, peak_src, kind = _conv_tensor_peak(heavy_meta["shape"], bf16_peak)
        roofline = {
            "kernel": f"{'conv_rs_kernel' if kind == 'f16' else 'conv_tc_kernel'} {heavy_meta['shape']} (B={B})", "bound": "tensor",
            "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf, "traffic": traffic,
            "traffic_note": (f"DRAM bytes/launch from ncu --set full at batch {traffic_batch} "
                             f"(profiles/r{'2' if kind == 'f16' else '1'}_traffic.json) x B/{traffic_batch}; " if traffic is not None else "") +
                            f"algorithmic bytes/launch = {heavy_meta['bytes']:.0f}",
            "peak_source": peak_src, "kernel_ms": kern_ms,
            "mma_tflops_executed": 3.0 * ach_tf, "mma_kind": kind, "mma_peak_tflops": 3.0 * peak_tf,
            "hbm_gbs_at_algorithmic_bytes": ach_gbs, "hbm_frac_at_algorithmic_bytes": ach_gbs / hbm_peak,
            "note": ("row-stationary implicit-GEMM conv on tcgen05 (kind::f16, fp16 hi / lo operand pairs from shared "
                     "memory, 3xFP16 for fp32-grade accuracy)" if kind == "f16" else
                     "implicit-GEMM conv on tcgen05 (kind::tf32, A via TMEM, 3xTF32 for fp32-grade accuracy)") +
                    "; achieved = algorithmic conv FLOPs / kernel time",
        }
    else:
        roofline = {
            "kernel": f"conv_kernel<3x3> {heavy_meta['shape']} (B={B})", "bound": "hbm",
            "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak, "traffic": None,
            "peak_source": which, "kernel_ms": kern_ms,
            "note": "fp32 FFMA direct conv: math-pipe bound (K=9*Cin per output), HBM fraction at algorithmic bytes "
                    "is necessarily small; see flop_* keys",
            "flop_pipe": "fp32_ffma", "flop_achieved_tflops": ach_tf, "flop_peak_tflops": FP32_FFMA_PEAK_TFLOPS,
            "flop_frac": ach_tf / FP32_FFMA_PEAK_TFLOPS,
        }
    roofline["step_hbm_frac_at_algorithmic_bytes"] = fps_gpu * bytes_frame / (hbm_peak * 1e9)
    roofline["step_algorithmic_tflops"] = fps_gpu * flops_frame / 1e12
    roofline["step_frac_of_fp32_ffma_peak"] = fps_gpu * flops_frame / (FP32_FFMA_PEAK_TFLOPS * 1e12)

    extra = {}
    del xs, host_x, host_u8, fe
    model._plans.clear()
    torch.cuda.empty_cache()
    if not args.no_other_configs:
        # BASELINE configs 2-4 as short device-timed sub-records (same kernels, same timing rules)
        extra["other_configs"] = run_other_configs(dev, world, rank, hbm_peak, bf16_peak, which)
    if not args.no_retrieval:
        from nano_vs_slam_b200 import retrieval_bench
        r = retrieval_bench.run(dev, world, rank)
        rf = r["roofline"]
        rf["peak"] = bf16_peak
        rf["frac"] = rf["achieved"] / bf16_peak
        rf["peak_source"] = which + " bf16 cuBLAS sustained (MEASURED_PEAKS.json); fp16 operands run at the bf16 rate"
        tr = _traffic("flat_l2_topk")
        rf["traffic"] = None
        if tr:
            n_local = retrieval_bench.shard_rows(world, rank)
            rf["traffic"] = tr["dram_bytes_per_launch"] * (n_local / tr["rows"] if "rows" in tr else 1.0)
            rf["traffic_note"] = tr["note"]
        if rank == 0 and not args.no_cpu_baseline:
            r["cpu_baseline"] = retrieval_cpu_baseline(os.cpu_count() or 1)
        extra["retrieval"] = r

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            fps_cpu, dt, kind, what = cpu_reference_fps(sd, 8, 4, threads, min_seconds=12.0)
            cpu = {"value": fps_cpu, "unit": "frames/s", "cores": threads, "kind": kind,
                   "sample": f"{round(fps_cpu * dt)} frames of the same workload in batches of 8 ({dt:.1f} s of CPU "
                             f"work), {what}"}
        line = {
            "metric": f"KP2DTiny-{LETTER} frames/s @{H}x{W}", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world, model.conv_backend),
            "e2e": dict({"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                         "ms_per_step": e2e_ms / args.steps}, **e2e_stats),
            "e2e_uint8_frames": e2e_u8,
            "gpu_launches": launches, "cuda_graph": bool(plan.graph is not None), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            # arithmetic of the tensor-core convs: "f16" = 3xFP16 row-stationary kernels (one MMA-issuing thread: results
            # are bit-reproducible run to run), "tf32" = 3xTF32 kernels (NVS_CONV_MATH)
            "conv_math": __import__("nano_vs_slam_b200.ops", fromlist=["conv_math"]).conv_math(),
        }
        line.update(extra)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
