#!/usr/bin/env python
"""Batch-axis attention at the sweep's size (B = 128, 1 327 packed positions, D = 768, H = 3) in
isolation: CUDA-event timing of forward / backward, and a stable launch sequence for ncu."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmu_b200 as mmu  # noqa: E402

PM = bool(int(os.environ.get("ATTN_PM", "1")))
B, L, D, H = 128, int(os.environ.get("ATTN_L", 1327)), 768, 3
qkv = (torch.randn(B * L, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dout = torch.randn(B * L, D, device="cuda").to(torch.bfloat16)


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out, probs = mmu.ops.attention_fwd(qkv, B, L, D, H, pos_major=PM)
if os.environ.get("ATTN_ONCE"):
    torch.cuda.synchronize()
    sys.exit(0)
us_f = timed(lambda: mmu.ops.attention_fwd(qkv, B, L, D, H, pos_major=PM))
us_b = timed(lambda: mmu.ops.attention_bwd(qkv, out, dout, probs, B, L, D, H, pos_major=PM), 5)
byts = L * H * 4 * B * (D // H) * 2
print(f"attention fwd {us_f:.1f} us ({byts / us_f / 1e6:.2f} TB/s of Q,K,V,O), bwd {us_b:.1f} us; L={L}")
