#!/usr/bin/env python
"""Which part of the train step breaks CUDA-graph capture?  python tools/debug_graph.py <stage>"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmu_b200 as mmu

stage = int(sys.argv[1])
mode = sys.argv[2] if len(sys.argv) > 2 else "global"
dev = torch.device("cuda")
B, E, C = 64, 4, 10
m = mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C).to(dev).train()
opt = torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9)
x = torch.rand(B, 4, 1, 14, 14, device=dev)
y = torch.randint(0, C, (B, E), device=dev)


def manual():
    saved = m._engine_forward(x, training=True)
    dl, _, _, accum = mmu.ops.heads_uncertainty_epilogue(saved[-1], y, 0, grad_scale=1.0 / (B * E), want_grad=True)
    m._engine_backward(saved, dl)
    opt.step()


def body():
    if stage == 7:
        return manual()
    if stage == 8:
        opt.zero_grad()
        logits, loss = m.forward_backward(x, y)
        opt.step()
        return mmu.acc(logits, y, False, True)
    if stage >= 6:
        opt.zero_grad()
    logits = m(x)
    if stage >= 2:
        loss = m.compute_loss(logits, y)
    if stage >= 3:
        loss.backward()
    if stage >= 4:
        opt.step()
    if stage >= 5:
        a = mmu.acc(logits, y, False, True)


body(); body()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, capture_error_mode=mode):
        body()
    g.replay()
    torch.cuda.synchronize()
    print(f"stage {stage} [{mode}]: captured and replayed")
except Exception as e:
    import traceback
    traceback.print_exc()
    print(f"stage {stage} [{mode}]: FAILED: {str(e)[:200]}".replace("\n", " | "))
