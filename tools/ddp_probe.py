#!/usr/bin/env python
"""Attribution of the data-parallel overhead of the headline training step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/ddp_probe.py [--steps 20]

For one process group it times the TRAIN step of bench.py's configuration (B = 128 per GPU,
197 + 40 tokens, D = 768, 3 layers, E = 5, C = 101, bf16, fused AdamW) under several gradient-
synchronisation set-ups, all with CUDA events (max over ranks):

  nocomm            gradients are not reduced at all: pure compute with all N GPUs busy
  r01               r01 behaviour: per-stage all-reduce on the default communicator, full GEMM grids
  serial_one_bucket ONE all-reduce of the whole flat gradient after the backward: nothing overlaps
  ctasK_reserveR    communicator capped at K CTAs, GEMM grids sized to (SMs - R) during the backward
  ..._mergeM        M backward stages per bucket

and records, per bucket, the duration of its all-reduce on the communication stream and the
duration of the backward on the compute stream.  Rank 0 prints one JSON document (commit it under
profiles/)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CFG = dict(B=128, l_img=197, l_txt=40, D=768, heads=3, layers=3, E=5, C=101)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import mmu_b200 as mmu
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)
    model = mmu.FlavaFusionTransfomer(out_dim=CFG["E"], num_classes=CFG["C"],
                                      multimodal_num_attention_heads=CFG["heads"],
                                      multimodal_num_hidden_layers=CFG["layers"], drop=0.0,
                                      avg_pool=False, precision="bf16").to(dev).train()
    opt = mmu.FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-3)
    g = torch.Generator().manual_seed(100 + rank)
    batches = []
    for _ in range(4):
        img = torch.randn(CFG["B"], CFG["l_img"], CFG["D"], generator=g).to(dev)
        txt = torch.randn(CFG["B"], CFG["l_txt"], CFG["D"], generator=g).to(dev)
        y = torch.randint(0, CFG["C"], (CFG["B"],), generator=g).to(dev)
        batches.append(((img, txt), y.unsqueeze(1).repeat(1, CFG["E"])))
    bwd_events = []

    def step(i, time_bwd=False):
        x, yt = batches[i % 4]
        opt.zero_grad()
        loss = model.compute_loss(model(x), yt)
        if time_bwd:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        loss.backward()
        if time_bwd:
            e1.record()
            bwd_events.append((e0, e1))
        opt.step()

    def timed(n):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i, time_bwd=True)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    configs = [("nocomm", None),
               ("r01", dict(nccl_max_ctas=0, reserve_sms=0, merge_stages=1)),
               ("serial_one_bucket", dict(nccl_max_ctas=0, reserve_sms=0, merge_stages=9)),
               ("ctas8_reserve0", dict(nccl_max_ctas=8, reserve_sms=0, merge_stages=1)),
               ("ctas8_reserve8", dict(nccl_max_ctas=8, reserve_sms=8, merge_stages=1)),
               ("ctas4_reserve4", dict(nccl_max_ctas=4, reserve_sms=4, merge_stages=1)),
               ("ctas16_reserve16", dict(nccl_max_ctas=16, reserve_sms=16, merge_stages=1)),
               ("ctas8_reserve8_merge2", dict(nccl_max_ctas=8, reserve_sms=8, merge_stages=2)),
               ("ctas8_reserve8_merge5", dict(nccl_max_ctas=8, reserve_sms=8, merge_stages=5)),
               ("ctas2_reserve2", dict(nccl_max_ctas=2, reserve_sms=2, merge_stages=1))]
    report = {"world": world, "steps": args.steps, "config": CFG, "results": {}}
    # the evaluation sweep involves no collective at all: its time at N GPUs against N = 1 shows
    # what part of the 1 -> N slowdown is NOT communication (shared power budget, host contention)
    torch.manual_seed(0)
    variants = [mmu.robustness.mask_level_variant(CFG["l_img"], CFG["l_txt"], "image", k, 10) for k in range(10)]
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model.forward_variants(batches[0][0], variants)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            model.forward_variants(batches[i % 4][0], variants)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        report["sweep10_ms_no_collective"] = round(float(t), 3)
    model.train()
    for name, kw in configs:
        model._ddp = None
        opt.grad_scale = 1.0
        ddp = None
        try:
            if kw is not None:
                ddp = mmu.parallel.DataParallel(model, opt, record_events=True, **kw)
            for i in range(args.warmup):
                step(i)
            bwd_events.clear()
            ms = timed(args.steps)
            bwd = sorted(a.elapsed_time(b) for a, b in bwd_events)
            entry = {"ms_per_train_step": round(ms, 3), "backward_ms_median": round(bwd[len(bwd) // 2], 3)}
            if ddp is not None:
                entry["buckets_last_step"] = [{"last_stage": st, "MB": round(nb / 1e6, 2), "allreduce_ms": round(t, 3),
                                               "GBps_algo": round(nb / 1e6 / max(t, 1e-6), 1)}
                                              for st, nb, t in ddp.bucket_times()]
                entry["allreduce_ms_sum"] = round(sum(b["allreduce_ms"] for b in entry["buckets_last_step"]), 3)
            report["results"][name] = entry
        except Exception as e:  # noqa: BLE001 -- one set-up failing must not lose the others
            report["results"][name] = {"error": repr(e)[:300]}
        torch.cuda.synchronize()
        dist.barrier()
    model._ddp = None
    base = report["results"].get("nocomm", {}).get("ms_per_train_step")
    if base:
        for name, r in report["results"].items():
            if "ms_per_train_step" in r:
                r["exposed_ms_vs_nocomm"] = round(r["ms_per_train_step"] - base, 3)
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
