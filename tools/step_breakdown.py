"""Per-launch breakdown of one forward step with CUDA events (warm, no profiler): every plan step is timed in
turn over a few repetitions of the whole step.  python tools/step_breakdown.py --batch 256 [--letter S --v3]"""
import argparse, contextlib, io, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nano_vs_slam_b200 import ops, tiny_factory
from nano_vs_slam_b200.synthetic import spread_init, synthetic_frames
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--letter", default="S")
ap.add_argument("--v3", action="store_true")
ap.add_argument("--classes", type=int, default=28)
ap.add_argument("--height", type=int, default=240)
ap.add_argument("--width", type=int, default=320)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
with contextlib.redirect_stdout(io.StringIO()):
    m = tiny_factory(a.letter, a.classes, v3=a.v3)
m.load_state_dict(spread_init(m.state_dict(), 1234)); m.eval(); m.training = False; m = m.cuda()
m.cuda_graph_max_batch = 0
x = synthetic_frames(a.batch, a.height, a.width, 0).cuda()
def step():
    out = m(x); post = m.post_processing(out, a.height, a.width)
    return ops.select_keypoints(post["score"], post["coord"], post["feat"], 0.7, 1000)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record(); torch.cuda.synchronize()
total = e0.elapsed_time(e1) / 5
plan = next(iter(m._plans.values()))
rows = []
for i, st in enumerate(plan.steps):
    plan.profile = {"idx": i, "events": []}
    for _ in range(a.reps):
        step()
    torch.cuda.synchronize()
    ms = statistics.median(s.elapsed_time(e) for s, e in plan.profile["events"])
    meta = plan.meta.get(i, {})
    rows.append((i, st[0], ms, meta))
plan.profile = None
fwd = sum(r[2] for r in rows)
print(f"step {total:.3f} ms ({a.batch / total * 1e3:.0f} frames/s); sum of forward launches {fwd:.3f} ms")
for i, kind, ms, meta in rows:
    tf = meta.get("flops", 0) / ms / 1e9 if ms > 0 else 0
    gbs = meta.get("bytes", 0) / ms / 1e6 if ms > 0 else 0
    print(f"{i:3d} {kind:5s} {ms:8.4f} ms {100 * ms / total:5.1f}%  {tf:7.1f} TF/s {gbs:7.0f} GB/s  {meta.get('shape', '')}")
