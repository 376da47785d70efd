#!/usr/bin/env python
"""Benchmark of the hot path: Food-101-shaped late fusion over synthetic FLAVA embeddings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload headline|sweep10|sweep43]

Headline workload (default) -- BASELINE.json configs[1] with configs[2]'s per-batch sweep.
One STEP on a batch of B = 128 samples per GPU (197 image + 40 text tokens of 768 features):
  1. robustness sweep of the batch over 10 mask levels (token-subset gather, eval forward of every
     level packed along the token axis into one pass -- all FLOPs of all levels executed --, fused
     uncertainty / ECE-histogram epilogue);
  2. GUIDED MODALITY DROPOUT (configs[1]): per sample, with probability p = 0.25, the modality the
     model is currently more confident about on its own (confidence of the image-only level vs the
     text-only level of step 1, still on the device) is zero-filled -- mmu_modality_keep_mask;
  3. training step on the masked batch: forward, CE over the 5 heads, backward, fused AdamW
     (cosine warm-up schedule), `acc`.
`value` = samples per second through that step (each sample is swept over 10 levels and trained on
once), whole job over N GPUs (weak scaling: B per GPU fixed).  Rank 0 prints ONE JSON line.

* `value`   : inputs resident in HBM when the timed region starts (CUDA events, max over ranks)
* `e2e`     : same step through the public API from pinned HOST buffers: `DevicePrefetcher`
              (H2D of batch i+1 on a side stream while step i runs) -> `forward_variants` +
              `UncertaintyMeter` -> `Model_.train_step(keep_mask=...)` -> D2H read of loss / acc;
              every batch is copied host->device inside the timed region
* `roofline`: dominant kernel = gemm_bf16_tcgen05_kernel (tensor bound); achieved = algorithmic
              FLOPs of one transformer block's 12 fwd+bwd GEMM launches / their CUDA-event time,
              timed live in a loop of >= 1 s -> the SUSTAINED measured peak is the denominator
              (`frac_of_burst` is printed next to it)
* `whole_step`: algorithmic FLOPs of everything the step computes / the step time
* `incumbent`: the reference's forward/backward expressed with fused ATen ops (oracle/eager.py,
              pinned to the reference goldens) run by PyTorch EAGER ON THE SAME B200 -- cuBLAS +
              ATen, fp32 as the reference runs it and bf16-autocast + fused AdamW as the strongest
              stock configuration: the real bar
* `cpu_baseline` / `--impl reference`: the oracle port (the reference is pure Python + torch CPU
              and cannot travel to the GPU box) on the host cores, the FULL B = 128 step.
* `--workload sweep10`: configs[2] alone -- sweep-only over `steps x B x N` pairs x 10 levels with
  the accumulator all-reduce inside the timed region (`--steps 977 --gpus 8` = 1 M pairs);
  `--workload sweep43`: configs[4] -- the reference's 43-variant schedule
  (eval_transformer_robustness.py:99-125) at 197 + 40 tokens, one meter per variant.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=128, l_img=197, l_txt=40, D=768, heads=3, layers=3, E=5, C=101, levels=10,
           lr=1e-3, wd=1e-3, p_drop=0.25)
WORKLOAD_NAME = ("configs[1] Food-101-shaped FLAVA late fusion, 5 heads, 101 classes, guided modality "
                 "dropout + configs[2] per-batch 10-level mask sweep")


# ----------------------------------------------------------------------------- helpers
def fwd_flops_per_sample(n_img, n_txt, B=None):
    """BASELINE.md 4.6 / SURVEY 8d: 2 L D^2 (projections) + layers (24 L D^2 + 4 L B D) + 2 E D C."""
    B = B or CFG["B"]
    D, L = CFG["D"], n_img + n_txt
    return 2 * L * D * D + CFG["layers"] * (24 * L * D * D + 4 * L * B * D) + 2 * CFG["E"] * D * CFG["C"]


def level_token_counts():
    out = []
    for k in range(CFG["levels"]):
        n = int(round(k * CFG["l_img"] / (CFG["levels"] - 1)))
        out.append((min(n, CFG["l_img"]), min(CFG["l_img"] - n, CFG["l_txt"])))
    return out


def step_flops(B=None):
    """Algorithmic FLOPs of one headline step per GPU: train (fwd + dgrad + wgrad = 3x fwd, the two
    input projections have no dgrad) + one eval forward per sweep level, all positions as written."""
    B = B or CFG["B"]
    D = CFG["D"]
    L = CFG["l_img"] + CFG["l_txt"]
    train = 3 * fwd_flops_per_sample(CFG["l_img"], CFG["l_txt"], B) - 2 * L * D * D
    sweep = sum(fwd_flops_per_sample(a, b, B) for a, b in level_token_counts())
    return B * (train + sweep), B * train, B * sweep


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"],
                    tensor_sustained=p["bf16_tflops_sustained"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0,
                source="fallback (B200_PROFILING.md)")


def gemm_traffic():
    """Average dram__bytes_read+write per GEMM launch from the committed ncu capture (or None)."""
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["avg_dram_bytes_per_launch"]
        except Exception:
            continue
    return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.window = [None, None]  # host-clock bounds of the timed region

    def start(self):
        """Starts sampling (every 20 ms); call mark_begin()/mark_end() around the timed region.
        The sampler is started before the warm-up so the process is already streaming when the
        short timed region begins; only samples taken inside the region are reported (all
        under-load samples if the region was too short to catch three)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.window[0] = time.time()

    def mark_end(self):
        self.window[1] = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        t0, t1 = self.window
        inside = [r[1:] for r in self.rows if t0 is not None and t1 is not None and t0 <= r[0] <= t1]
        scope = "timed region"
        if len(inside) < 3:
            inside = [r[1:] for r in self.rows if t0 is None or r[0] >= t0 - 2.0]
            scope = "warm-up + timed region (region shorter than 3 samples)"
        self.rows = inside
        self.scope = scope
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        smax = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "scope": self.scope}


def make_host_batches(n, B, seed, pin, dtype=torch.float32):
    out = []
    for i in range(n):
        g = torch.Generator().manual_seed(seed + i)
        img = torch.randn(B, CFG["l_img"], CFG["D"], generator=g).to(dtype)
        txt = torch.randn(B, CFG["l_txt"], CFG["D"], generator=g).to(dtype)
        y = torch.randint(0, CFG["C"], (B,), generator=g)
        if pin:
            img, txt, y = img.pin_memory(), txt.pin_memory(), y.pin_memory()
        out.append(((img, txt), y))
    return out


def draw_level_variants(mask_level_variant, seed):
    """10 mask levels of the image modality, host RNG, drawn per batch: level 0 is
    text-only (40 tokens), level 9 image-only (197 tokens)."""
    # the CPU generator only (what torch.randperm draws from): torch.manual_seed also queues a seed
    # call, with a formatted stack trace, for every uninitialised accelerator backend -- 0.35 ms
    torch.default_generator.manual_seed(seed)
    return [mask_level_variant(CFG["l_img"], CFG["l_txt"], "image", k, CFG["levels"])
            for k in range(CFG["levels"])]


def reference_init_state_dict(seed=42):
    """Initial weights exactly as the reference constructor draws them (torch's own module
    constructors in the reference's order, src/model.py:225-256) -- plain torch, no product code:
    the CPU / incumbent arms must not load the library they are compared against."""
    import torch.nn as nn
    torch.manual_seed(seed)
    D, sd = CFG["D"], {}
    for i in range(CFG["layers"]):
        pre = f"mm_encoder.resblocks.{i}."
        attn = nn.MultiheadAttention(D, CFG["heads"])
        sd[pre + "attn.in_proj_weight"], sd[pre + "attn.in_proj_bias"] = attn.in_proj_weight, attn.in_proj_bias
        sd[pre + "attn.out_proj.weight"], sd[pre + "attn.out_proj.bias"] = attn.out_proj.weight, attn.out_proj.bias
        ln1 = nn.LayerNorm(D)
        fc, proj = nn.Linear(D, 4 * D), nn.Linear(4 * D, D)
        ln2 = nn.LayerNorm(D)
        for k, m in (("ln_1", ln1), ("mlp.c_fc", fc), ("mlp.c_proj", proj), ("ln_2", ln2)):
            sd[pre + k + ".weight"], sd[pre + k + ".bias"] = m.weight, m.bias
    for k, m in (("ln_pre", nn.LayerNorm(D)), ("ln_post", nn.LayerNorm(D)),
                 ("image_to_mm_projection", nn.Linear(CFG["D"], D)),
                 ("text_to_mm_projection", nn.Linear(CFG["D"], D))):
        sd[k + ".weight"], sd[k + ".bias"] = m.weight, m.bias
    for e in range(CFG["E"]):
        m = nn.Linear(D, CFG["C"])
        sd[f"output_layers.{e}.weight"], sd[f"output_layers.{e}.bias"] = m.weight, m.bias
    return {k: v.detach().clone() for k, v in sd.items()}


# ------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    import mmu_b200 as mmu
    from functools import partial

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = CFG["B"]

    torch.manual_seed(42)
    model = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"],
                                      multimodal_num_attention_heads=CFG["heads"],
                                      multimodal_num_hidden_layers=CFG["layers"], drop=0.0,
                                      avg_pool=False, precision="bf16",
                                      live_tokens=args.live_tokens)
    opt = mmu.FusedAdamW(model.parameters(), lr=CFG["lr"], betas=(0.9, 0.98), eps=1e-9,
                         weight_decay=CFG["wd"])
    sched = mmu.get_cosine_schedule_with_warmup(opt, 3 * 100, 100 * 100)
    shaping = partial(mmu.dataset.data_forming_func_transformer, model_type="MultiHead")

    def multihead5(x, y, phase):  # config[1] has 5 heads: tile the labels over all of them
        x, y = shaping(x, y, phase)
        return x, (y[:, :1].repeat(1, CFG["E"]) if phase == "train" else y)

    trainer = mmu.Model_(model, opt, sched, multihead5, metrics=[mmu.acc], verbose=False)
    trainer.to(dev)
    ddp = mmu.parallel.DataParallel(model, opt) if world > 1 else None  # noqa: F841 (hooks the model)
    meter = mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"])

    nb = 4
    # --host-dtype bf16: the embeddings are staged on the host in bf16 (half the host->device bytes;
    # the bf16 engine rounds its inputs to bf16 in the stem anyway, so results are bit-identical to
    # feeding the fp32 values that round to them).  Default fp32: the reference's stored format.
    host = make_host_batches(nb, B, 1000 * (rank + 1), pin=True,
                             dtype=torch.bfloat16 if args.host_dtype == "bf16" else torch.float32)
    resident = [((i.to(dev), t.to(dev)), y.to(dev)) for (i, t), y in host]
    y_rep = {}

    def sweep_and_mask(img, txt, y, seed):
        """Steps 1 + 2: packed 10-level sweep + fused epilogue (per-sample scores stay on the
        device), then the guided keep mask from the image-only / text-only confidences."""
        variants = draw_level_variants(mmu.robustness.mask_level_variant, seed)
        model.eval()
        with torch.no_grad():
            logits = model.forward_variants((img, txt), variants)        # (levels, B, E, C)
            _, scores = meter.update(logits.view(-1, CFG["E"], CFG["C"]), y.repeat(len(variants)),
                                     want_scores=True)                    # (levels*B, 4)
        model.train()
        lv = CFG["levels"]
        return mmu.robustness.modality_dropout_mask_device(
            B, CFG["p_drop"], "guided", dev, score_img=scores[(lv - 1) * B:, 0], score_txt=scores[:B, 0])

    def step_resident(i):
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

    def step_e2e(batch, i):
        (img, txt), y = batch          # device tensors from the prefetcher (copied this step)
        keep = sweep_and_mask(img, txt, y, i)
        return trainer.train_step((img, txt), y, keep_mask=keep, sync=False)[:2]   # device scalars

    def run_e2e(steps):
        """Public-API loop: pinned host batches -> DevicePrefetcher (H2D of batch i+1 on a side
        stream while step i runs) -> packed sweep -> Model_.train_step.  Every step's batch is
        copied host->device and every step's loss / acc is read back device->host inside the timed
        region; the read of step i is collected after step i+1 has been enqueued (one-step lag,
        the way an asynchronous logger consumes them) and is copied on a side stream, so the host
        never drains the queue."""
        prefetcher.loader = [host[i % nb] for i in range(steps)]
        pending, history = None, []
        for i, batch in enumerate(prefetcher):
            loss, info = step_e2e(batch, i)
            # D2H of this step's loss / acc on a side stream behind an event (metrics.AsyncScalars):
            # a plain float(loss) would copy on the compute stream, i.e. behind the step that has
            # just been enqueued, and stall the host until THAT step ends
            ticket = reader.push([loss] + list(info))
            if pending is not None:
                history.append(reader.pop(pending))
            pending = ticket
        history.append(reader.pop(pending))
        return history

    prefetcher = mmu.dataset.DevicePrefetcher([], dev)  # its two device slots persist across runs
    reader = mmu.metrics.AsyncScalars(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.mark_begin()
        n0 = mmu._lib.lib.mmu_launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        if sampler:
            sampler.mark_end()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        launches = mmu._lib.lib.mmu_launch_count() - n0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, clocks

    model.train()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    if args.workload != "headline":
        line = run_sweep_workload(args, mmu, model, dev, resident, nb, timed, sampler, world, rank)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    for i in range(args.warmup):
        step_resident(i)
    # Both timed legs should see the board in the SAME state.  A B200 under this step reaches its
    # power cap after a few hundred ms; the second leg used to start on an already capped board --
    # an apparent 3-6 % "end-to-end overhead" that an interleaved measurement (tools/e2e_probe.py:
    # 13.47 vs 13.43-13.52 ms) and the CUPTI timeline of the e2e loop (0.05 ms idle per step) do
    # not show.  So each leg now starts the same way: >= --leg-pause-s of idle, then W warm-up steps.
    # --settle-steps N (default 0) adds N untimed steps before the first leg instead, for a
    # sustained-state number (same count on every rank; reported as config.settle_steps).
    settle_steps = 0 if args.profile else max(0, int(args.settle_steps))
    for i in range(settle_steps):
        step_resident(i)
    ms, launches, clocks = timed(step_resident, args.steps, sampler)
    meter.all_reduce()
    summary = meter.compute()
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms / args.steps,
                              "gpu_launches": int(launches)}))
        return
    if args.leg_pause_s > 0:
        barrier()
        time.sleep(args.leg_pause_s)
    run_e2e(max(2, args.warmup))
    meter.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_e2e(args.steps)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t)
    meter.all_reduce()
    meter.compute()  # D2H read of the sweep result
    per_step = ms / args.steps
    value = world * B / (per_step / 1e3)
    e2e_value = world * B / (ms_e2e / args.steps / 1e3)

    # ---- rooflines, measured live
    roof = cpu = hbm_kernels = other = incumbent = None
    pk = peaks()
    if rank == 0:
        roof, hbm_kernels = measure_rooflines(mmu, dev)
        if world == 1:
            if not args.no_incumbent:
                del resident
                torch.cuda.empty_cache()
                incumbent = incumbent_eager(dev)
            if not args.no_cpu:
                cpu = cpu_reference(steps=1, warmup=0)
            if not args.no_other_configs:
                other = other_configs(mmu, dev)
    if world > 1:
        dist.barrier()

    if rank == 0:
        total, train_f, sweep_f = step_flops()
        live = bool(args.live_tokens)
        h2d = world * sum(t.numel() * t.element_size() for t in (host[0][0][0], host[0][0][1], host[0][1]))
        h2d += world * 2 * B * 4   # the two uniform vectors of the dropout mask
        whole = None
        if not live:
            tf = total / (per_step * 1e-3) / 1e12
            whole = {"bound": "tensor", "achieved": round(tf, 1), "peak": pk["tensor_sustained"],
                     "unit": "TFLOP/s", "frac": round(tf / pk["tensor_sustained"], 3),
                     "flops_per_step": total, "train_flops": train_f, "sweep_flops": sweep_f,
                     "note": "algorithmic FLOPs of the whole step (all token positions as written, "
                             "BASELINE.md 4.6) / step time; sustained measured peak"}
        line = {
            "metric": "train+robustness-eval samples/sec", "value": round(value, 2),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME,
                       "per_gpu_batch": B, "global_batch": B * world, "l_img": CFG["l_img"],
                       "l_txt": CFG["l_txt"], "D": CFG["D"], "layers": CFG["layers"],
                       "heads": CFG["heads"], "E": CFG["E"], "C": CFG["C"],
                       "mask_levels": CFG["levels"], "optimizer": "fused AdamW",
                       "modality_dropout": f"guided, p={CFG['p_drop']}",
                       "parallelism": f"dp{world}",
                       "dead_tokens": "skipped (live-token path: only positions < E computed)" if live
                       else "computed (as written)",
                       "host_dtype": args.host_dtype,
                       "legs": f"value, then e2e; each after idle (>= {args.leg_pause_s} s before e2e) + warm-up steps; "
                               f"{settle_steps} extra settle steps before value",
                       "l2": "working set ~6 GB/step >> 126 MB L2, inputs rotate over 4 batches"},
            "e2e": {"value": round(e2e_value, 2), "unit": "samples/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * world,  # loss + acc, 4 B each, per rank
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "whole_step": whole,
            "hbm_kernels": hbm_kernels,
            "cpu_baseline": cpu,
            "incumbent": incumbent,
            "sweep_summary": {k: summary[k] for k in ("acc", "ece", "h_pred", "mi", "n_samples")},
            "other_configs": other,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sweep_workload(args, mmu, model, dev, resident, nb, timed, sampler, world, rank):
    """configs[2] (`sweep10`) / configs[4] (`sweep43`) alone: robustness evaluation only.  Every rank
    sweeps its own `steps` batches of B pairs (pool of 4 resident batches, index sets drawn per
    batch from the host RNG); the integer histogram / metric accumulators are sum-all-reduced
    INSIDE the timed region after the last batch (bit-exact merge)."""
    B = CFG["B"]
    model.eval()
    if args.workload == "sweep10":
        meters = [mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"])]
        n_var = CFG["levels"]

        def variants_for(i):
            return draw_level_variants(mmu.robustness.mask_level_variant, i)
    else:
        import numpy as np
        n_var = 43
        meters = [mmu.metrics.UncertaintyMeter(dev, CFG["C"], CFG["E"]) for _ in range(n_var)]

        def variants_for(i):
            np.random.seed(i)
            torch.manual_seed(i)
            return mmu.robustness.robustness_variants(CFG["l_img"], CFG["l_txt"], 20)
    positions = [0]

    def step(i):
        (img, txt), y = resident[i % nb]
        variants = variants_for(i)
        positions[0] = sum((len(a) if a is not None else 0) + (len(b) if b is not None else 0)
                           for a, b in variants)
        with torch.no_grad():
            logits = model.forward_variants((img, txt), variants)
            if len(meters) == 1:
                meters[0].update(logits.view(-1, CFG["E"], CFG["C"]), y.repeat(len(variants)))
            else:
                for lg, m in zip(logits, meters):
                    m.update(lg, y)

    def run(steps):
        def body(i):
            step(i)
            if i == steps - 1:
                for m in meters:
                    m.all_reduce()
        return body

    for i in range(args.warmup):
        step(i)
    for m in meters:
        m.reset()
    ms, launches, clocks = timed(run(args.steps), args.steps, sampler)
    summ = meters[0].compute()
    per_step = ms / args.steps
    flops = B * sum(fwd_flops_per_sample(a, b) for a, b in level_token_counts()) if args.workload == "sweep10" else None
    pk = peaks()
    line = {"metric": "robustness-eval samples/sec", "value": round(world * B / (per_step / 1e3), 2),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("configs[2] robustness sweep: pairs x 10 mask levels" if args.workload == "sweep10"
                                    else "configs[4] 43-variant schedule of eval_transformer_robustness.py:99-125"),
                       "pairs_total": args.steps * B * world, "variants_per_pair": n_var,
                       "token_positions_per_pair": positions[0], "per_gpu_batch": B, "l_img": CFG["l_img"],
                       "l_txt": CFG["l_txt"], "E": CFG["E"], "C": CFG["C"], "parallelism": f"dp{world}",
                       "collective": "sum-all-reduce of the metric accumulators, inside the timed region",
                       "dead_tokens": "skipped" if args.live_tokens else "computed (as written)",
                       "l2": "inputs rotate over 4 resident batches (93 MB each)"},
            "variant_evals_per_s": round(world * B * n_var / (per_step / 1e3), 1),
            "gpu_launches": int(launches), "clocks": clocks,
            "sweep_summary": {k: summ[k] for k in ("acc", "ece", "h_pred", "mi", "n_samples")}}
    if flops and not args.live_tokens:
        tf = flops / (per_step * 1e-3) / 1e12
        line["whole_step"] = {"bound": "tensor", "achieved": round(tf, 1), "peak": pk["tensor_sustained"],
                              "unit": "TFLOP/s", "frac": round(tf / pk["tensor_sustained"], 3)}
    return line


def other_configs(mmu, dev):
    """Short device-timed measurements (CUDA events, 3 warm-up + 5 steps) of the BASELINE.json configs
    the headline does not cover -- reported next to it, never part of `value`: configs[3] MMBT
    (BERT-base + ResNet-152 image tokens, 512 positions, batch 32, bf16: train step from pooled
    tokens and from raw images, robustness forward) and configs[0] (four-view FashionMNIST ResNet,
    batch 256).  N = 1 only; any failure is reported as a string instead of breaking the line."""
    import types
    out = {}

    def timed(fn, n=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        Bm, S_txt, n_img = 32, 507, 3
        vocab = types.SimpleNamespace(stoi={"[CLS]": 101, "[SEP]": 102, "[PAD]": 0})
        g = torch.Generator().manual_seed(42)
        txt = torch.randint(1000, 30522, (Bm, S_txt), generator=g)
        lens = torch.randint(S_txt // 2, S_txt + 1, (Bm,), generator=g)
        mask = (torch.arange(S_txt)[None] < lens[:, None]).long()
        txt, segment = (txt * mask).to(dev), mask.clone().to(dev)
        mask = mask.to(dev)
        y = torch.randint(0, 2, (Bm,), generator=g).to(dev)
        S, D, L = n_img + 2 + S_txt, 768, 12
        flops_fwd = L * (24 * S * D * D + 4 * S * S * D) * Bm
        for images in (False, True):
            args = types.SimpleNamespace(bert_model="bert-base-uncased", hidden_sz=768, img_hidden_sz=2048,
                                         num_image_embeds=n_img, img_embed_pool_type="avg", dropout=0.0,
                                         n_classes=2, vocab=vocab, precision="bf16",
                                         img_encoder="native" if images else None,
                                         bert_dropout=0.0)
            torch.manual_seed(42)
            m = mmu.MultimodalBertClf(args).to(dev).train()
            named = list(m.named_parameters())
            nd = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
            opt = mmu.BertAdam([{"params": [p for n, p in named if not any(k in n for k in nd)], "weight_decay": 0.01},
                                {"params": [p for n, p in named if any(k in n for k in nd)], "weight_decay": 0.0}],
                               lr=5e-5, warmup=0.1, t_total=1000)
            img = (torch.randn(Bm, 3, 224, 224, generator=g) if images
                   else torch.randn(Bm, n_img, 2048, generator=g)).to(dev)

            def train_step():
                opt.zero_grad()
                loss = m.compute_loss(m(txt, mask, segment, img), y)
                loss.backward()
                opt.step()

            ms = timed(train_step)
            key = "mmbt_images" if images else "mmbt_tokens"
            out[key] = {"train_ms": round(ms, 2), "train_samples_per_s": round(Bm / ms * 1e3, 1)}
            if not images:
                out[key]["train_tflops_trunk"] = round(3 * flops_fwd / ms / 1e9, 1)
            m.eval()
            with torch.no_grad():
                ms = timed(lambda: m(txt, mask, segment, img))
            out[key].update(eval_forward_ms=round(ms, 2), eval_samples_per_s=round(Bm / ms * 1e3, 1))
            del m, opt
            torch.cuda.empty_cache()
        out["mmbt_config"] = "BERT-base, 3 + 2 + 507 = 512 positions, batch 32, bf16, BertAdam, binary head"
    except Exception as e:  # noqa: BLE001 -- a side measurement must never break the headline line
        out["mmbt_error"] = repr(e)[:200]
    try:
        Bf, E, C = 256, 4, 10
        g = torch.Generator().manual_seed(42)
        x = torch.rand(Bf, 4, 1, 14, 14, generator=g).to(dev)
        yt = torch.randint(0, C, (Bf,), generator=g).unsqueeze(1).repeat(1, E).to(dev)
        for prec in ("fp32", "bf16"):
            torch.manual_seed(42)
            m = mmu.MIMOResNet(num_channels=1, emb_dim=4, out_dim=E, num_classes=C, precision=prec).to(dev).train()
            opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)

            def step():
                opt.zero_grad()
                m.compute_loss(m(x), yt).backward()
                opt.step()

            ms = timed(step, 10)
            out["fmnist_mimo_resnet_" + prec] = {"train_ms": round(ms, 3), "train_samples_per_s": round(Bf / ms * 1e3, 1)}
            # the same step replayed from a CUDA graph through the trainer (graphs.GraphedTrainStep)
            tr = mmu.Model_(m, opt, None, lambda a, b, phase="train": (a, b), metrics=[mmu.acc], verbose=False)
            tr.to(dev)
            ms = timed(lambda: tr.train_step(x, yt, sync=False, cuda_graph=True), 10)
            out["fmnist_mimo_resnet_" + prec].update(graph_train_ms=round(ms, 3),
                                                     graph_train_samples_per_s=round(Bf / ms * 1e3, 1))
    except Exception as e:  # noqa: BLE001
        out["fmnist_error"] = repr(e)[:200]
    try:
        # post-hoc rank statistics (exact pair counting, compute bound): AUROC of 43 variants of a
        # 10 000-sample evaluation set in one batched launch, and one 200 000-score vector
        g = torch.Generator().manual_seed(42)
        lab = torch.randint(0, 2, (10000,), generator=g).float().to(dev)
        sc = torch.randn(43, 10000, generator=g).to(dev)
        ms = timed(lambda: mmu.ops.pair_concordance(lab, sc))
        big_l = torch.randint(0, 2, (200000,), generator=g).float().to(dev)
        big_s = torch.randn(200000, generator=g).to(dev)
        ms2 = timed(lambda: mmu.ops.pair_concordance(big_l, big_s), 3)
        out["rank_stats"] = {"auroc_43x10k_ms": round(ms, 3), "pairs_per_s_43x10k": round(43 * 10000 * 9999 / 2 / ms * 1e3),
                             "auroc_200k_ms": round(ms2, 3), "pairs_per_s_200k": round(200000 * 199999 / 2 / ms2 * 1e3)}
    except Exception as e:  # noqa: BLE001
        out["rank_stats_error"] = repr(e)[:200]
    return out


def measure_rooflines(mmu, dev, min_seconds=1.0):
    """Times, with CUDA events on the launching stream, (a) the GEMM launches of one transformer
    block's forward + backward at the headline size, looped for >= `min_seconds` so that the board
    is in its sustained (power-capped) state -- the denominator is therefore the SUSTAINED measured
    cuBLAS bf16 peak, and the burst fraction is printed beside it -- and (b) the two HBM-bound
    kernels."""
    pk = peaks()
    B, L, D = CFG["B"], CFG["l_img"] + CFG["l_txt"], CFG["D"]
    M = B * L
    E_ = mmu._lib
    bf = torch.bfloat16
    x768 = torch.randn(M, D, device=dev).to(bf)
    x2304 = torch.randn(M, 3 * D, device=dev).to(bf)
    x3072 = torch.randn(M, 4 * D, device=dev).to(bf)
    w = {n: (torch.randn(n, D, device=dev) * 0.02).to(bf) for n in (3 * D, 4 * D, D)}  # keyed by out-features
    w_proj = (torch.randn(D, 4 * D, device=dev) * 0.02).to(bf)
    o2304, o3072, o3072b, o768 = (torch.empty(M, n, device=dev, dtype=bf) for n in (3 * D, 4 * D, 4 * D, D))
    gw = {s: torch.zeros(*s, device=dev) for s in ((3 * D, D), (4 * D, D), (D, 4 * D), (D, D))}
    bias = {n: torch.zeros(n, device=dev) for n in (D, 3 * D, 4 * D)}
    g = mmu.ops.gemm

    # ---- HBM-bound kernels FIRST (each timed alone in a short loop on a cool board: the measured copy
    #      bandwidth is the peak), before the >= 1 s GEMM loop drives the board into its power cap
    hbm = []
    n = 22_843_392 // 4 * 4
    p, gr, m_, v_ = (torch.randn(n, device=dev) for _ in range(4))
    v_.abs_()
    sh = torch.empty(n, device=dev, dtype=bf)
    for _ in range(3):
        mmu.ops.adamw_flat_step(p, gr, m_, v_, 1, 1e-3, p_bf16=sh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        mmu.ops.adamw_flat_step(p, gr, m_, v_, i + 2, 1e-3, p_bf16=sh)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    gbs = 28.0 * n / (us * 1e-6) / 1e9
    hbm.append({"kernel": "adamw_kernel", "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"],
                "unit": "GB/s", "frac": round(gbs / pk["hbm"], 3), "bytes_per_param": 28,
                "us_per_launch": round(us, 1), "note": "22.8 M params (640 MB/launch of p,g,m,v; > L2)"})
    N = 1 << 20
    logits = torch.randn(N, CFG["E"], CFG["C"], device=dev)
    y = torch.randint(0, CFG["C"], (N,), device=dev)
    acc = mmu.ops.new_accum(dev)
    for _ in range(2):
        mmu.ops.heads_uncertainty_epilogue(logits, y, 1, accum=acc)
    e0.record()
    for _ in range(5):
        mmu.ops.heads_uncertainty_epilogue(logits, y, 1, accum=acc)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    byts = N * (CFG["E"] * CFG["C"] * 4 + 8)
    gbs = byts / (us * 1e-6) / 1e9
    hbm.append({"kernel": "ce_uncertainty_kernel", "bound": "hbm", "achieved": round(gbs, 1),
                "peak": pk["hbm"], "unit": "GB/s", "frac": round(gbs / pk["hbm"], 3),
                "bytes_per_sample": CFG["E"] * CFG["C"] * 4 + 8, "us_per_launch": round(us, 1),
                "note": "1 Mi samples x (5 x 101) logits, eval mode (2.1 GB/launch; > L2)"})

    def layer_gemms():
        # forward
        g(x768, w[3 * D], out=o2304, bias=bias[3 * D])
        g(x768, w[D], out=o768, bias=bias[D])
        g(x768, w[4 * D], mode=E_.EPI_QUICKGELU, out=o3072, out2=o3072b, bias=bias[4 * D])
        g(x3072, w_proj, out=o768, bias=bias[D])
        # backward: dgrad
        g(x768, w_proj, b_mn_major=True, mode=E_.EPI_DGELU, out=o3072, aux=o3072b)
        g(x3072, w[4 * D], b_mn_major=True, out=o768)
        g(x768, w[D], b_mn_major=True, out=o768)
        g(x2304, w[3 * D], b_mn_major=True, out=o768)
        # backward: wgrad (split-K atomics)
        g(x768, x3072, a_mn_major=True, b_mn_major=True, mode=E_.EPI_ATOMIC, out=gw[(D, 4 * D)], splits=8)
        g(x3072, x768, a_mn_major=True, b_mn_major=True, mode=E_.EPI_ATOMIC, out=gw[(4 * D, D)], splits=2)
        g(x768, x768, a_mn_major=True, b_mn_major=True, mode=E_.EPI_ATOMIC, out=gw[(D, D)], splits=8)
        g(x2304, x768, a_mn_major=True, b_mn_major=True, mode=E_.EPI_ATOMIC, out=gw[(3 * D, D)], splits=3)

    flops_layer = 3 * (2 * M * D * 3 * D + 2 * M * D * D + 2 * 2 * M * D * 4 * D)
    for _ in range(3):
        layer_gemms()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        layer_gemms()
    e1.record()
    torch.cuda.synchronize()
    burst_ms = e0.elapsed_time(e1) / 10
    reps = max(20, int(min_seconds * 1e3 / burst_ms) + 1)
    n0 = mmu._lib.lib.mmu_launch_count()
    e0.record()
    for _ in range(reps):
        layer_gemms()
    e1.record()
    torch.cuda.synchronize()
    n_launch = mmu._lib.lib.mmu_launch_count() - n0
    total_ms = e0.elapsed_time(e1)
    ms = total_ms / reps
    tf = flops_layer / (ms * 1e-3) / 1e12
    tf_burst = flops_layer / (burst_ms * 1e-3) / 1e12
    roof = {"kernel": "gemm_bf16_tcgen05_kernel", "bound": "tensor", "achieved": round(tf, 1),
            "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "frac": round(tf / pk["tensor_sustained"], 3),
            "timed_seconds": round(total_ms / 1e3, 2),
            "burst": {"achieved": round(tf_burst, 1), "peak": pk["tensor_burst"],
                      "frac": round(tf_burst / pk["tensor_burst"], 3),
                      "note": "first 10 repetitions (~10 ms) against the burst peak"},
            "peak_source": pk["source"] + ": sustained cuBLAS bf16 for the >= 1 s loop, burst for the 10 ms loop",
            "launches_per_rep": int(n_launch // reps), "avg_launch_us": round(ms * 1e3 / (n_launch / reps), 1),
            "flops_per_launch_avg": flops_layer / (n_launch / reps), "traffic": gemm_traffic(),
            "note": "12 GEMM launches of one transformer block's fwd+bwd (M=30336), operands > L2; "
                    "traffic = average DRAM bytes per launch from the committed ncu capture under profiles/"}

    # ---- the sweep's GEMMs: the four launches of one block's EVAL forward at the packed sweep's size
    #      (in_proj / c_fc with the LayerNorm folded into the epilogue, out_proj / c_proj with the
    #      residual-stream epilogue), in the same sustained state
    del x768, x2304, x3072, o2304, o3072, o3072b, o768
    Ms = B * sum(a + b for a, b in level_token_counts())
    nt = (D + 127) // 128
    raw = torch.randn(Ms, D, device=dev).to(bf)
    xs = torch.randn(Ms, D, device=dev)
    qkv = torch.empty(Ms, 3 * D, device=dev, dtype=bf)
    u = torch.empty(Ms, 4 * D, device=dev, dtype=bf)
    stats = torch.stack([xs.sum(1), (xs * xs).sum(1)], 1).reshape(Ms, 1, 2).repeat(1, nt, 1).contiguous() / nt
    cw = {n: torch.zeros(n, device=dev) for n in (3 * D, 4 * D)}
    xs2, raw2, stats2 = torch.empty_like(xs), torch.empty_like(raw), torch.empty_like(stats)

    def eval_gemms():   # outputs go to separate buffers so that repetitions see the same inputs
        g(raw, w[3 * D], out=qkv, bias=bias[3 * D], ln_fold=(stats, cw[3 * D], 1e-5))
        g(raw, w[D], mode=E_.EPI_RESID_LN, out=xs2, out2=raw2, bias=bias[D], aux=xs, stats_out=stats2)
        g(raw, w[4 * D], mode=E_.EPI_QUICKGELU, out2=u, bias=bias[4 * D], ln_fold=(stats, cw[4 * D], 1e-5))
        g(u, w_proj, mode=E_.EPI_RESID_LN, out=xs2, out2=raw2, bias=bias[D], aux=xs, stats_out=stats2)

    flops_eval = 2 * Ms * D * 3 * D + 2 * Ms * D * D + 2 * 2 * Ms * D * 4 * D
    for _ in range(3):
        eval_gemms()
    torch.cuda.synchronize()
    reps_e = max(20, int(0.5 * min_seconds * 1e3 / (flops_eval / 1.1e15 * 1e3)) + 1)
    e0.record()
    for _ in range(reps_e):
        eval_gemms()
    e1.record()
    torch.cuda.synchronize()
    ms_e = e0.elapsed_time(e1) / reps_e
    tf_e = flops_eval / (ms_e * 1e-3) / 1e12
    roof["eval_gemms"] = {
        "achieved": round(tf_e, 1), "peak": pk["tensor_sustained"], "unit": "TFLOP/s",
        "frac": round(tf_e / pk["tensor_sustained"], 3), "rows": Ms, "launches_per_rep": 4,
        "ms_per_rep": round(ms_e, 3), "timed_seconds": round(ms_e * reps_e / 1e3, 2),
        "note": "one block's eval forward at the packed 10-level sweep's size: in_proj / c_fc with ln_1 / ln_2 "
                "folded into the epilogue, out_proj / c_proj adding the fp32 residual stream in theirs "
                "(10 B/element of residual traffic ride inside these launches)"}
    return roof, hbm


# ---------------------------------------------------------------------- incumbent / CPU arms
def _oracle_step_factory(device, trainer):
    """The headline step in the reference's own terms (one forward per sweep level, autograd,
    torch.optim.AdamW), shared by the eager-on-GPU incumbent and described by oracle/eager.py."""
    from oracle import shaping, uncertainty

    def step(i, img, txt, y):
        variants = draw_level_variants(shaping.mask_level_variant, i)
        logits = trainer.sweep((img, txt), variants)                       # (levels, B, E, C)
        s_img = uncertainty.ensemble_scores(logits[-1])["conf"]
        s_txt = uncertainty.ensemble_scores(logits[0])["conf"]
        keep = shaping.modality_dropout_mask(img.shape[0], CFG["p_drop"], "guided",
                                             torch.stack([s_img, s_txt], 1).cpu())
        m_img, m_txt = shaping.apply_keep_mask(img, txt, keep.to(device))
        yt = y.unsqueeze(1).repeat(1, CFG["E"])
        return trainer.train_step((m_img, m_txt), yt)
    return step


def incumbent_eager(dev):
    """PyTorch EAGER on the same B200 (cuBLAS / ATen / SDPA kernels; none of this repo's kernels):
    the headline step with the reference's arithmetic (oracle/eager.py, pinned to the reference
    goldens by tests/test_oracle_golden.py).  Two configurations: fp32 exactly as the reference
    runs, and bf16 autocast + fused torch AdamW (the strongest stock setting).  Note the guided mask
    is built on the host here (one D2H read per step), as a stock implementation would."""
    from oracle import eager
    out = {}
    (img, txt), y = make_host_batches(1, CFG["B"], 7, pin=False)[0]
    img, txt, y = img.to(dev), txt.to(dev), y.to(dev)
    for name, kw in (("fp32_as_reference", dict()),
                     ("bf16_autocast_fused_adamw", dict(fused_optimizer=True, autocast=torch.bfloat16))):
        try:
            P = {k: v.to(dev) for k, v in reference_init_state_dict().items()}
            tr = eager.EagerTrainer(P, CFG["heads"], CFG["layers"], CFG["E"], CFG["lr"], CFG["wd"], **kw)
            step = _oracle_step_factory(dev, tr)
            for i in range(3):
                step(i, img, txt, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for i in range(n):
                step(i, img, txt, y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = {"value": round(CFG["B"] / ms * 1e3, 1), "unit": "samples/s", "ms_per_step": round(ms, 2)}
            del tr, P
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001 -- a baseline leg must never break the headline line
            out[name] = {"error": repr(e)[:200]}
    out["kind"] = ("PyTorch 2.11 eager on the same B200, same step, B=128: oracle/eager.py (fused ATen "
                   "ops of the reference's nn.Modules), 3 warm-up + 5 timed steps, CUDA events")
    return out


def cpu_reference(steps, warmup, B=None):
    """The oracle port of the reference's CPU path (torch CPU, fp32, all host threads): the FULL
    headline step (B = 128: 10-level sweep, guided mask, train step, AdamW).  No product code is
    imported on this path.  Returns the cpu_baseline object."""
    from oracle import fusion, optim, shaping, uncertainty
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    B = B or CFG["B"]
    P = reference_init_state_dict()
    m = {k: torch.zeros_like(v) for k, v in P.items()}
    v = {k: torch.zeros_like(vv) for k, vv in P.items()}
    (img, txt), y = make_host_batches(1, B, 7, pin=False)[0]
    yt = y.unsqueeze(1).repeat(1, CFG["E"])

    def step(i):
        nonlocal P, m, v
        variants = draw_level_variants(shaping.mask_level_variant, i)
        confs = {}
        with torch.no_grad():
            for k, var in enumerate(variants):
                s_img, s_txt = shaping.apply_variant(img, txt, var)
                lg = fusion.flava_fusion_forward(P, (s_img, s_txt), CFG["heads"], False)
                uncertainty.calibration_histograms(lg, y)
                if k in (0, CFG["levels"] - 1):
                    confs[k] = uncertainty.ensemble_scores(lg)["conf"]
        keep = shaping.modality_dropout_mask(B, CFG["p_drop"], "guided",
                                             torch.stack([confs[CFG["levels"] - 1], confs[0]], 1))
        m_img, m_txt = shaping.apply_keep_mask(img, txt, keep)
        logits, loss, grads = fusion.loss_and_grads(P, (m_img, m_txt), yt, CFG["heads"], False)
        for k in P:
            P[k], m[k], v[k] = optim.adamw_step(P[k], grads[k], m[k], v[k], i + 1, CFG["lr"])
        fusion.acc(logits, yt, False, True)

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(B / dt, 3), "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"the full step, B={B} of {CFG['B']}, {steps} timed step(s) after {warmup} warm-up; "
                      f"torch-CPU fp32 oracle port of the reference modules (the reference itself is "
                      f"Python and cannot travel to the GPU box)",
            "s_per_step": round(dt, 3)}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    cpu = cpu_reference(steps, warm)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "train+robustness-eval samples/sec", "value": cpu["value"],
            "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": cpu["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME,
                       "per_gpu_batch": CFG["B"], "global_batch": CFG["B"], "l_img": CFG["l_img"],
                       "l_txt": CFG["l_txt"], "D": CFG["D"], "layers": CFG["layers"],
                       "heads": CFG["heads"], "E": CFG["E"], "C": CFG["C"],
                       "mask_levels": CFG["levels"], "modality_dropout": f"guided, p={CFG['p_drop']}",
                       "note": "CPU arm: rank 0 only, one process, the full B=128 step"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=["headline", "sweep10", "sweep43"])
    ap.add_argument("--live-tokens", action="store_true",
                    help="skip the token positions that cannot reach the logits (avg_pool=False: only "
                         "positions < E are live, src/model.py:286-287); bit-identical logits")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short side measurements of configs[0] / configs[3] (N = 1 only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the inline cpu_baseline leg")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the eager-on-B200 incumbent leg")
    ap.add_argument("--rooflines-only", action="store_true",
                    help="only the per-kernel roofline legs (for ncu --metrics dram__bytes_* captures)")
    ap.add_argument("--host-dtype", choices=("fp32", "bf16"), default="fp32",
                    help="dtype of the host-side embeddings (bf16: half the H2D bytes, bit-identical results "
                         "for values that are bf16-representable)")
    ap.add_argument("--settle-steps", type=int, default=0,
                    help="untimed extra steps beyond --warmup before the first timed leg (40 = ~0.5 s: the board "
                         "is in its sustained, power-capped state when the timing starts)")
    ap.add_argument("--leg-pause-s", type=float, default=1.5,
                    help="idle time before the warm-up of the second (e2e) timed leg, so that it starts from the "
                         "same board state as the first")
    ap.add_argument("--profile", action="store_true",
                    help="skip the e2e / roofline / CPU legs (short run for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
        if args.rooflines_only:
            import mmu_b200 as mmu
            torch.cuda.set_device(0)
            roof, hbm = measure_rooflines(mmu, torch.device("cuda", 0), min_seconds=0.05)
            print(json.dumps({"roofline": roof, "hbm_kernels": hbm}))
            return
        run_gpu(args)


if __name__ == "__main__":
    main()
