#!/usr/bin/env python
"""One batched AUROC (43 variants x 10 000 samples) and one 200 000-score pair count, for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmu_b200 as mmu
dev = torch.device("cuda")
g = torch.Generator().manual_seed(42)
lab = torch.randint(0, 2, (10000,), generator=g).float().to(dev)
sc = torch.randn(43, 10000, generator=g).to(dev)
big_l = torch.randint(0, 2, (200000,), generator=g).float().to(dev)
big_s = torch.randn(200000, generator=g).to(dev)
for _ in range(2):
    mmu.ops.pair_concordance(lab, sc)
    mmu.ops.pair_concordance(big_l, big_s)
torch.cuda.synchronize()
