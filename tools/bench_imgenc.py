#!/usr/bin/env python
"""MMBT image encoder alone (reference src/mmbt.py:15-45: ResNet-152 trunk + adaptive pool), batch 32
of 224x224 images, bf16 tensor-core convolutions: eval forward and train forward+backward, CUDA
events.  Algorithmic FLOPs: 2 x 11.5 G MACs per image forward (torchvision resnet152), x3 for training."""
import argparse, json, os, sys, types
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib
import mmu_b200 as mmu
ie = importlib.import_module("multi-modal-uncertainty_b200.src.image_encoder")
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda:0")
args = types.SimpleNamespace(num_image_embeds=3, img_embed_pool_type="avg", precision=a.precision)
torch.manual_seed(0)
enc = ie.ImageEncoder(args).to(dev)
x = torch.randn(a.batch, 3, 224, 224, device=dev)
r = torch.randn(a.batch, 3, 2048, device=dev)


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    l0 = mmu._lib.lib.mmu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (mmu._lib.lib.mmu_launch_count() - l0) // n


def train():
    enc.zero_grad()
    (enc(x) * r).sum().backward()


flops = 2 * 11.5e9 * a.batch
out = {"config": {"workload": "ImageEncoder resnet152 trunk, batch %d x 3 x 224 x 224, %s" % (a.batch, a.precision)}}
enc.train()
ms, nl = timed(train, a.steps)
out["train_fwd_bwd"] = {"ms": round(ms, 2), "images_per_s": round(a.batch / ms * 1e3, 1), "tflops": round(3 * flops / ms / 1e9, 1), "gpu_launches": int(nl)}
enc.eval()
with torch.no_grad():
    ms, nl = timed(lambda: enc(x), a.steps)
out["eval_fwd"] = {"ms": round(ms, 2), "images_per_s": round(a.batch / ms * 1e3, 1), "tflops": round(flops / ms / 1e9, 1), "gpu_launches": int(nl)}
print(json.dumps(out))
