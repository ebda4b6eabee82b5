"""Measured dense TF32 tensor peak of this pool's B200 ("tf32 peak is derived as bf16 / 2, never measured"):
torch.matmul on fp32 operands with TF32 allowed (cuBLAS tf32 tensor-core GEMM), 8192^3, burst (best of 10) and sustained
(back to back for 4 s) -- the recipe MEASURED_PEAKS.json uses for bf16.  Writes one JSON line."""
import json
import time

import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = time.time(); it = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    it += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / it
# bf16 the same way, same box, for the ratio
ah, bh = a.bfloat16(), b.bfloat16()
for _ in range(3):
    ah @ bh
torch.cuda.synchronize()
e0.record()
for _ in range(200):
    ah @ bh
e1.record(); torch.cuda.synchronize()
bf = e0.elapsed_time(e1) / 200
fl = 2.0 * n ** 3
print(json.dumps({"tf32_tflops": fl / best / 1e9, "tf32_tflops_sustained": fl / sus / 1e9,
                  "bf16_tflops_same_box_200_iters": fl / bf / 1e9, "gpu": torch.cuda.get_device_name(0),
                  "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS): best of 10, and back to back for 4 s"}))
