#!/usr/bin/env python
"""BASELINE.json configs[3]: MMBT, BERT-base text + 3 pooled ResNet-152 image tokens, sequence
3 + 2 + 507 = 512, batch 32 per GPU, bf16 operands, binary head (hateful-memes shaped).  Train step
(forward + backward + fused BertAdam) and the forward_control robustness forward, timed with CUDA
events; algorithmic FLOPs = 3 x 12 x (24 S D^2 + 4 S^2 D) per sample.  `--tokens`
(default) feeds pooled image tokens (frozen / cached image encoder); `--images` runs the CUDA
ResNet-152 image encoder inside the step.  The CPU oracle port is timed on a bounded sample."""
import argparse, json, os, sys, time, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mmu_b200 as mmu

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--s-txt", type=int, default=507)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--images", action="store_true")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--bert-dropout", type=float, default=0.0, help="BERT hidden / attention dropout (reference config: 0.1)")
a = ap.parse_args()
# one process per GPU under torchrun (weak scaling: `--batch` samples per rank, no data-path
# collective but the gradient all-reduce inside optimizer.step())
RANK, WORLD = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
LOCAL = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
dev = torch.device("cuda", LOCAL)
B, S_txt, n_img, C = a.batch, a.s_txt, 3, 2
vocab = types.SimpleNamespace(stoi={"[CLS]": 101, "[SEP]": 102, "[PAD]": 0})
args = types.SimpleNamespace(bert_model="bert-base-uncased", hidden_sz=768, img_hidden_sz=2048,
                             num_image_embeds=n_img, img_embed_pool_type="avg", dropout=0.0, n_classes=C,
                             vocab=vocab, precision=a.precision, img_encoder="native" if a.images else None,
                             bert_dropout=a.bert_dropout)
torch.manual_seed(42)
m = mmu.MultimodalBertClf(args).to(dev).train()
named = list(m.named_parameters())
no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
opt = mmu.BertAdam([{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
                    {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}],
                   lr=5e-5, warmup=0.1, t_total=1000)
if WORLD > 1:
    mmu.parallel.FlatGradSync.attach(opt)  # overlap=True measured slower (47.7 vs 40.1 ms from images): the all-reduce competes with the HBM-bound encoder backward
g = torch.Generator().manual_seed(42 + RANK)
txt = torch.randint(1000, 30522, (B, S_txt), generator=g)
lens = torch.randint(S_txt // 2, S_txt + 1, (B,), generator=g)
mask = (torch.arange(S_txt)[None] < lens[:, None]).long()
txt, segment = txt * mask, mask.clone()
img = torch.randn(B, 3, 224, 224, generator=g) if a.images else torch.randn(B, n_img, 2048, generator=g)
y = torch.randint(0, C, (B,), generator=g)
txt, mask, segment, img, y = (t.to(dev) for t in (txt, mask, segment, img, y))


def train_step():
    opt.zero_grad()
    logits = m(txt, mask, segment, img)
    loss = m.compute_loss(logits, y)
    loss.backward()
    opt.step()
    return loss


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    l0 = mmu._lib.lib.mmu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if WORLD > 1:  # max over ranks, on the device clock
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms, r, (mmu._lib.lib.mmu_launch_count() - l0) // n


S, D, L = n_img + 2 + S_txt, 768, 12
flops_fwd = L * (24 * S * D * D + 4 * S * S * D) * B
out = {"config": {"workload": "MMBT bert-base, seq %d, batch %d, %s, %s" % (S, B, a.precision, "images" if a.images else "pooled image tokens")}}
ms, loss, nl = timed(train_step, a.steps)
out["n_gpus"] = WORLD
out["train"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(WORLD * B / ms * 1e3, 1), "loss": float(loss.detach()),
                "tflops": round(3 * flops_fwd / ms / 1e9, 1), "gpu_launches": int(nl)}
m.eval()
with torch.no_grad():
    ms, _, nl = timed(lambda: m(txt, mask, segment, img), a.steps)
    out["eval_forward"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(WORLD * B / ms * 1e3, 1),
                           "tflops": round(flops_fwd / ms / 1e9, 1), "gpu_launches": int(nl)}
    torch.manual_seed(0)
    ms, _, nl = timed(lambda: m.forward_control(txt, mask, segment, img, "text"), a.steps)
    out["forward_control_text"] = {"ms_per_step": round(ms, 3), "samples_per_s": round(WORLD * B / ms * 1e3, 1)}

    # the whole robustness sweep of one batch (eval_mmbt_robustness.py:76-96): full + img_only + txt_only +
    # 20 + 20 forward_control draws = 43 variants; image tokens computed once per batch
    torch.manual_seed(0)
    batch = [((txt, mask, segment, img), y)]
    ms, _, nl = timed(lambda: mmu.robustness.run_mmbt_robustness(m, batch, n_repeats=20, device=dev), max(2, a.steps // 3))
    out["robustness_sweep_43"] = {"ms_per_batch": round(ms, 2), "samples_per_s": round(WORLD * B / ms * 1e3, 1),
                                  "variant_forwards_per_s": round(WORLD * 43 * B / ms * 1e3, 1), "gpu_launches": int(nl)}

if not a.no_cpu and not a.images and RANK == 0:
    # CPU oracle port, fp32, all host threads, bounded sample: batch 2 of the same shape, 1 train step
    from oracle import mmbt as O
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    P = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    cfg = dict(n_layers=12, n_head=12, cls_id=101, sep_id=102)
    sl = slice(0, 2)
    t0 = time.perf_counter()
    O.loss_and_grads(P, txt[sl].cpu(), mask[sl].cpu(), segment[sl].cpu(), img[sl].cpu(), y[sl].cpu(), cfg)
    dt = time.perf_counter() - t0
    out["cpu_port_train"] = {"s_per_step": round(dt, 3), "samples_per_s": round(2 / dt, 2), "cores": cores,
                             "sample": "batch 2 of the same sequence length, 1 step (oracle/mmbt.py, fp32)"}
if RANK == 0:
    print(json.dumps(out))
if WORLD > 1:
    dist.destroy_process_group()
