"""Pinned host -> device and device -> pinned host copy rates at the bench's batch sizes."""
import torch
x = torch.empty(256, 3, 240, 320).pin_memory()
d = torch.empty_like(x, device="cuda")
o = torch.empty(79356928 // 4, device="cuda")
h = torch.empty(79356928 // 4).pin_memory()
for name, fn, nbytes in (("H2D 236 MB", lambda: d.copy_(x, non_blocking=True), x.numel() * 4),
                         ("D2H 79 MB", lambda: h.copy_(o, non_blocking=True), o.numel() * 4)):
    ts = []
    for _ in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "ms:", " ".join(f"{t:.2f}" for t in ts), f"-> best {nbytes / min(ts) / 1e6:.1f} GB/s, worst {nbytes / max(ts[2:]) / 1e6:.1f} GB/s")
