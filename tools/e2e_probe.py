#!/usr/bin/env python
"""Where do the 0.7 ms between bench.py's `value` (resident inputs) and `e2e` (pinned host ->
prefetcher -> public API -> D2H of loss / acc) go?  Times the headline step under switches:

  resident_manual   : bench.py's step_resident (the `value` leg)
  resident_trainer  : resident batches through Model_.train_step(sync=False), no read-back
  prefetch_noread   : pinned host batches through DevicePrefetcher, no read-back
  prefetch_read_lag : + float() of the previous step's loss / acc (bench.py's `e2e` leg)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mmu_b200 as mmu  # noqa: E402
from functools import partial  # noqa: E402

CFG = bench.CFG
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(42)
model = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"], multimodal_num_attention_heads=CFG["heads"],
                                  multimodal_num_hidden_layers=CFG["layers"], drop=0.0, avg_pool=False, precision="bf16")
opt = mmu.FusedAdamW(model.parameters(), lr=CFG["lr"], betas=(0.9, 0.98), eps=1e-9, weight_decay=CFG["wd"])
sched = mmu.get_cosine_schedule_with_warmup(opt, 300, 10000)
shaping = partial(mmu.dataset.data_forming_func_transformer, model_type="MultiHead")


def multihead5(x, y, phase):
    x, y = shaping(x, y, phase)
    return x, (y[:, :1].repeat(1, CFG["E"]) if phase == "train" else y)


trainer = mmu.Model_(model, opt, sched, multihead5, metrics=[mmu.acc], verbose=False)
trainer.to(dev)
meter = mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"])
B, nb = CFG["B"], 4
host = bench.make_host_batches(nb, B, 1000, pin=True)
resident = [((i.to(dev), t.to(dev)), y.to(dev)) for (i, t), y in host]


def sweep_and_mask(img, txt, y, seed):
    variants = bench.draw_level_variants(mmu.robustness.mask_level_variant, seed)
    model.eval()
    with torch.no_grad():
        logits = model.forward_variants((img, txt), variants)
        _, scores = meter.update(logits.view(-1, CFG["E"], CFG["C"]), y.repeat(len(variants)), want_scores=True)
    model.train()
    lv = CFG["levels"]
    return mmu.robustness.modality_dropout_mask_device(B, CFG["p_drop"], "guided", dev,
                                                       score_img=scores[(lv - 1) * B:, 0], score_txt=scores[:B, 0])


def step_manual(i):
    (img, txt), y = resident[i % nb]
    keep = sweep_and_mask(img, txt, y, i)
    yt = y.unsqueeze(1).repeat(1, CFG["E"])
    opt.zero_grad()
    logits = model((img, txt), keep_mask=keep)
    loss = model.compute_loss(logits, yt)
    loss.backward()
    opt.step()
    mmu.acc(logits, yt, False, True)
    sched.step()


def step_trainer(batch, i):
    (img, txt), y = batch
    keep = sweep_and_mask(img, txt, y, i)
    return trainer.train_step((img, txt), y, keep_mask=keep, sync=False)[:2]


def timed(fn, steps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_manual(steps):
    for i in range(steps):
        step_manual(i)


def run_resident_trainer(steps):
    for i in range(steps):
        step_trainer(resident[i % nb], i)


prefetcher = mmu.dataset.DevicePrefetcher([], dev)


def run_prefetch(steps, read):
    prefetcher.loader = [host[i % nb] for i in range(steps)]
    pending = None
    for i, batch in enumerate(prefetcher):
        cur = step_trainer(batch, i)
        if read and pending is not None:
            float(pending[0]), [float(v) for v in pending[1]]
        pending = cur
    if read:
        float(pending[0])


reader = mmu.metrics.AsyncScalars(dev)


def run_prefetch_async(steps):
    prefetcher.loader = [host[i % nb] for i in range(steps)]
    pending = None
    for i, batch in enumerate(prefetcher):
        loss, info = step_trainer(batch, i)
        ticket = reader.push([loss] + list(info))
        if pending is not None:
            reader.pop(pending)
        pending = ticket
    reader.pop(pending)


model.train()
legs = (("resident_manual", run_manual), ("resident_trainer", run_resident_trainer),
        ("prefetch_noread", lambda n: run_prefetch(n, False)),
        ("prefetch_read_item", lambda n: run_prefetch(n, True)),
        ("prefetch_read_async", run_prefetch_async))
for _, fn in legs:
    fn(4)
out = {name: [] for name, _ in legs}
for rnd in range(4):            # interleaved rounds cancel the thermal drift of the board
    for name, fn in legs:
        out[name].append(round(timed(fn, 10), 3))
out["mean"] = {k: round(sum(v) / len(v), 3) for k, v in out.items()}
print(json.dumps(out))

# ---- CUPTI view of the e2e loop: where does the GPU idle?  (MMU_E2E_TIMELINE=1)
if os.environ.get("MMU_E2E_TIMELINE"):
    from torch.profiler import ProfilerActivity, profile
    run_prefetch_async(30)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_prefetch_async(6)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    kern = sorted((e for e in ev if "Memcpy" not in e.name and "Memset" not in e.name), key=lambda e: e.time_range.start)
    copies = [e for e in ev if "Memcpy HtoD" in e.name]
    t0, t1 = kern[0].time_range.start, max(e.time_range.end for e in kern)
    gaps, cur = [], kern[0].time_range.end
    for e in kern[1:]:
        if e.time_range.start > cur + 5:      # > 5 us of no kernel running
            gaps.append((round(cur - t0, 1), round(e.time_range.start - cur, 1), e.name[:60]))
        cur = max(cur, e.time_range.end)
    gaps.sort(key=lambda g: -g[1])
    big = [e for e in copies if e.time_range.end - e.time_range.start > 100]
    print(json.dumps({"span_ms_per_step": (t1 - t0) / 6e3, "idle_ms_per_step": sum(g[1] for g in gaps) / 6e3,
                      "largest_gaps_us(at, len, next kernel)": gaps[:12],
                      "h2d_big_copies": len(big),
                      "h2d_ms_per_step": sum(e.time_range.end - e.time_range.start for e in big) / 6e3,
                      "h2d_GBps": [round((93194240 / 2 if False else 0), 1)][:0]}), file=sys.stderr)
