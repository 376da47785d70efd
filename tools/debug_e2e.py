#!/usr/bin/env python
"""Host-side timeline of the e2e step (where does the host block?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mmu_b200 as mmu
from functools import partial
CFG = bench.CFG
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"], avg_pool=False, precision="bf16")
opt = mmu.FusedAdamW(model.parameters(), lr=1e-3)
sched = mmu.get_cosine_schedule_with_warmup(opt, 300, 10000)
shaping = partial(mmu.dataset.data_forming_func_transformer, model_type="MultiHead")
def mh(x, y, phase):
    x, y = shaping(x, y, phase)
    return x, (y[:, :1].repeat(1, CFG["E"]) if phase == "train" else y)
trainer = mmu.Model_(model, opt, sched, mh, metrics=[mmu.acc], verbose=False).to(dev)
meter = mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"])
host = bench.make_host_batches(4, CFG["B"], 1000, pin=True)
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
model.train()
loader = [host[i % 4] for i in range(8)]
t_all = time.perf_counter()
for (img, txt), y in mmu.dataset.DevicePrefetcher(loader, dev):
    t0 = time.perf_counter(); loss, info, _ = trainer.train_step((img, txt), y, sync=False); tick("train_step", t0)
    t0 = time.perf_counter(); variants = bench.level_variants(mmu, 0); tick("variants", t0)
    model.eval()
    with torch.no_grad():
        t0 = time.perf_counter(); logits = model.forward_variants((img, txt), variants); tick("forward_variants", t0)
        t0 = time.perf_counter(); meter.update(logits.view(-1, CFG["E"], CFG["C"]), y.repeat(len(variants))); tick("meter", t0)
    model.train()
    t0 = time.perf_counter(); float(loss); tick("sync", t0)
torch.cuda.synchronize()
print("total ms/step", (time.perf_counter() - t_all) * 1e3 / 8)
for k, v in T.items():
    print(f"{k:18s}", " ".join(f"{x:7.2f}" for x in v))
