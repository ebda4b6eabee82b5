"""Attention core timing (config 4 shapes): python tools/bench_attention.py  [NVS_ATT_QPT=2|4]"""
import sys, torch
sys.path.insert(0, "/root/repo")
from nano_vs_slam_b200 import ops
for (B, C, h, w) in ((16, 64, 128, 256), (16, 64, 64, 128), (64, 64, 60, 80)):
    q = torch.randn(B, C, h, w, device="cuda")
    kv = torch.randn(B, 2 * C, h // 2, w // 2, device="cuda")
    for _ in range(2):
        out = ops.attention(q, kv, 4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = ops.attention(q, kv, 4)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4.0 * B * (h * w) * (h * w // 4) * C
    print(f"B={B} C={C} {h}x{w}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
