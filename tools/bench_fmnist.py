#!/usr/bin/env python
"""BASELINE.json configs[0]: four-view FashionMNIST, batch 256, fp32 -- train-step throughput of
MIMOResNet and MIMOTransfomer on the GPU next to the CPU oracle port (bounded sample)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmu_b200 as mmu

dev = torch.device("cuda:0")
B, E, C = 256, 4, 10
g = torch.Generator().manual_seed(42)
x = torch.rand(B, 4, 1, 14, 14, generator=g)
y = torch.randint(0, C, (B,), generator=g)
yt = y.unsqueeze(1).repeat(1, E)
out = {}
for name, make, opt_fn in (
        ("MIMOResNet", lambda: mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C),
         lambda m: torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)),
        ("MIMOResNet_bf16", lambda: mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C,
                                                   precision="bf16"),
         lambda m: torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)),
        ("MIMOTransfomer", lambda: mmu.MIMOTransfomer(out_dim=E, num_classes=C, hidden_size=768, precision="bf16"),
         lambda m: mmu.FusedAdamW(m.parameters(), lr=1e-3))):
    torch.manual_seed(42)
    m = make().to(dev).train()
    opt = opt_fn(m)
    xd, yd = x.to(dev), yt.to(dev)

    def step():
        opt.zero_grad()
        logits = m(xd)
        loss = m.compute_loss(logits, yd)
        loss.backward()
        opt.step()
        return loss

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    out[name] = {"ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1), "loss": float(loss.detach())}
    if isinstance(opt, torch.optim.SGD):
        # the same step through the reference-protocol trainer, replayed from a CUDA graph
        # (graphs.GraphedTrainStep); inputs come from pinned host memory every step
        tr = mmu.Model_(m, opt, None, lambda a, b, phase="train": (a, b), metrics=[mmu.acc], verbose=False)
        tr.to(dev)
        xh, yh = x.pin_memory(), yt.pin_memory()
        for mode in (False, True):
            for _ in range(5):
                tr.train_step(xh, yh, sync=False, cuda_graph=mode)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(30):
                l, info, _ = tr.train_step(xh, yh, sync=False, cuda_graph=mode)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 30
            out[name + ("_trainer_graph" if mode else "_trainer_eager")] = {
                "ms_per_step": round(ms, 3), "samples_per_s": round(B / ms * 1e3, 1), "loss": float(l),
                "acc": float(info[0])}

# CPU oracle port of the ResNet train step (fp32, all host threads), bounded: 2 steps
from oracle import resnet
cores = len(os.sched_getaffinity(0))
torch.set_num_threads(cores)
torch.manual_seed(42)
P = {k: v.detach().clone() for k, v in mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C).state_dict().items()}
resnet.loss_and_grads(P, x, yt, C)
t0 = time.perf_counter()
for _ in range(2):
    resnet.loss_and_grads(P, x, yt, C)
dt = (time.perf_counter() - t0) / 2
out["cpu_port_MIMOResNet"] = {"s_per_step": round(dt, 3), "samples_per_s": round(B / dt, 1), "cores": cores}
print(json.dumps(out))
