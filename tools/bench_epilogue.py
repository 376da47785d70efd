#!/usr/bin/env python
"""Micro-benchmark of ce_uncertainty_kernel (eval and train+gradient modes) against the HBM roofline."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmu_b200 as mmu

dev = torch.device("cuda:0")
E, C = int(os.environ.get("E", 5)), int(os.environ.get("C", 101))
N = int(os.environ.get("N", 1 << 20))
reps = int(os.environ.get("REPS", 5))
logits = torch.randn(N, E, C, device=dev)
y = torch.randint(0, C, (N,), device=dev)
yt = y.unsqueeze(1).repeat(1, E).contiguous()
acc = mmu.ops.new_accum(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
def eval_hb1():
    os.environ["MMU_CE_HB"] = "1"
    mmu.ops.heads_uncertainty_epilogue(logits, y, 1, accum=acc)
    del os.environ["MMU_CE_HB"]
for name, fn, byts in (
        ("eval_head_by_head", eval_hb1, N * (E * C * 4 + 8)),
        ("eval", lambda: mmu.ops.heads_uncertainty_epilogue(logits, y, 1, accum=acc), N * (E * C * 4 + 8)),
        ("train_grad", lambda: mmu.ops.heads_uncertainty_epilogue(logits, yt, 0, grad_scale=1.0 / (N * E),
                                                                  want_grad=True, accum=acc),
         N * (2 * E * C * 4 + 8 * E))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    out[name] = {"us": round(us, 1), "GBps": round(byts / us / 1e3, 1)}
print(json.dumps(out))
