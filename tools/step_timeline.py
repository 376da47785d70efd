#!/usr/bin/env python
"""In-situ kernel timeline of the headline step (bench.py's `value` leg): runs a few steps under
torch.profiler (CUPTI activity records: concurrent, warm, power-capped -- unlike ncu's serialised
cold-cache replay) and prints, per kernel name, total / mean duration and its share of the span,
plus the span's idle time.  usage: step_timeline.py [steps] > out.json"""
import json
import os
import re
import sys
from functools import partial

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mmu_b200 as mmu  # noqa: E402

CFG = bench.CFG
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
torch.manual_seed(42)
model = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"], multimodal_num_attention_heads=CFG["heads"],
                                  multimodal_num_hidden_layers=CFG["layers"], drop=0.0, avg_pool=False,
                                  precision="bf16", live_tokens=bool(os.environ.get("MMU_TL_LIVE"))).to(dev)
opt = mmu.FusedAdamW(model.parameters(), lr=CFG["lr"], betas=(0.9, 0.98), eps=1e-9, weight_decay=CFG["wd"])
sched = mmu.get_cosine_schedule_with_warmup(opt, 300, 10000)
meter = mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"])
B, nb = CFG["B"], 4
host = bench.make_host_batches(nb, B, 1000, pin=True)
resident = [((i.to(dev), t.to(dev)), y.to(dev)) for (i, t), y in host]


def step(i):
    (img, txt), y = resident[i % nb]
    variants = bench.draw_level_variants(mmu.robustness.mask_level_variant, i)
    model.eval()
    with torch.no_grad():
        logits = model.forward_variants((img, txt), variants)
        _, scores = meter.update(logits.view(-1, CFG["E"], CFG["C"]), y.repeat(len(variants)), want_scores=True)
    model.train()
    lv = CFG["levels"]
    keep = mmu.robustness.modality_dropout_mask_device(B, CFG["p_drop"], "guided", dev,
                                                       score_img=scores[(lv - 1) * B:, 0], score_txt=scores[:B, 0])
    yt = y.unsqueeze(1).repeat(1, CFG["E"])
    opt.zero_grad()
    out = model((img, txt), keep_mask=keep)
    loss = model.compute_loss(out, yt)
    loss.backward()
    opt.step()
    mmu.acc(out, yt, False, True)
    sched.step()


for i in range(30):     # long enough for the power cap to engage, as in the bench
    step(i)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy, cur_end = 0.0, t0
for e in ev:                      # union of kernel intervals (kernels may overlap under PDL)
    s, t = e.time_range.start, e.time_range.end
    if t > cur_end:
        busy += t - max(s, cur_end)
        cur_end = t
agg = {}
for e in ev:
    name = e.name.replace("mmu::<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    a = agg.setdefault(name, [0.0, 0])
    a[0] += e.time_range.end - e.time_range.start
    a[1] += 1
span = t1 - t0
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
out = {"steps": steps, "span_ms_per_step": span / steps / 1e3, "busy_ms_per_step": busy / steps / 1e3,
       "idle_frac": 1 - busy / span,
       "kernels": [{"name": k[:110], "ms_per_step": v[0] / steps / 1e3, "n_per_step": v[1] / steps,
                    "us_each": v[0] / v[1], "share": v[0] / span} for k, v in rows[:40]]}
print(json.dumps(out, indent=1))

# host enqueue time per step (no synchronisation inside the loop): is the step host-bound?
import time  # noqa: E402
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(20):
    step(i)
host_ms = (time.perf_counter() - t) / 20 * 1e3
torch.cuda.synchronize()
total_ms = (time.perf_counter() - t) / 20 * 1e3
print(json.dumps({"host_enqueue_ms_per_step": host_ms, "wall_ms_per_step": total_ms}), file=sys.stderr)
