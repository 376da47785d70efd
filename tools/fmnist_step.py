#!/usr/bin/env python
"""Three FashionMNIST MIMOResNet train steps (batch 256) for an ncu launch list: PREC=fp32|bf16."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmu_b200 as mmu
dev = torch.device("cuda")
m = mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=4, num_classes=10,
                   precision=os.environ.get("PREC", "bf16")).to(dev).train()
opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
x = torch.rand(256, 4, 1, 14, 14, device=dev)
y = torch.randint(0, 10, (256, 4), device=dev)
for _ in range(3):
    opt.zero_grad()
    m.forward_backward(x, y)
    opt.step()
torch.cuda.synchronize()
