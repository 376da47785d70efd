"""``ImageEncoder`` of the MMBT path (reference ``src/mmbt.py:15-45``): torchvision's ResNet-152
trunk (``children()[:-2]``) + adaptive pooling to ``num_image_embeds`` tokens, executed by the CUDA
engine in ``csrc/resnet.cu`` (im2col + GEMM convolutions -- tcgen05 with ``precision='bf16'`` for
every convolution but the 3-channel stem --, BatchNorm with batch / running statistics, max pool,
adaptive avg / max pool, full backward).

State-dict keys are the reference's (``model.0.weight``, ``model.1.running_mean``,
``model.4.0.conv1.weight``, ``model.7.2.bn3.num_batches_tracked`` ...), so a reference checkpoint's
``enc.img_encoder.*`` entries load with ``strict=True``.  Parameters are views of one flat fp32
buffer, running statistics of a second one.  There is no CPU path and no pretrained download: the
weights are random-initialised (torchvision's scheme) until a checkpoint is loaded.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from ._backend import _lib
from .model import FlavaFusionTransfomer, _holder_for, check_workspace, stamp_workspace
from .resnet import MIMOResNet

#: reference src/mmbt.py:28-37: num_image_embeds -> adaptive pool grid
POOL_GRID = {1: (1, 1), 2: (2, 1), 3: (3, 1), 5: (5, 1), 7: (7, 1), 4: (2, 2), 6: (3, 2), 8: (4, 2), 9: (3, 3)}


class _EncoderForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, enc, x):
        ctx.enc = enc
        ctx.saved = enc._engine_forward(x, training=True)
        return ctx.saved[-1]

    @staticmethod
    def backward(ctx, dtokens):
        ctx.enc._engine_backward(ctx.saved, dtokens.contiguous())
        return None, None, None


class ImageEncoder(nn.Module):
    """Drop-in for reference ``ImageEncoder(args)``; ``forward(x)``: ``(B, 3, H, H)`` ->
    ``(B, num_image_embeds, 2048)``.  Optional ``args.img_encoder_layers`` (default resnet152's
    ``(3, 8, 36, 3)``), ``args.img_encoder_width`` (torchvision ``width_per_group``, default 64) and
    ``args.precision`` ("bf16" default / "fp32")."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.precision = getattr(args, "precision", "bf16")
        self._layers = tuple(int(v) for v in getattr(args, "img_encoder_layers", (3, 8, 36, 3)))
        self._width = int(getattr(args, "img_encoder_width", 64))
        if args.num_image_embeds not in POOL_GRID:
            raise ValueError("num_image_embeds must be one of 1..9 (src/mmbt.py:28-37)")
        self._pool = POOL_GRID[int(args.num_image_embeds)]
        self._pool_max = int(args.img_embed_pool_type != "avg")
        self._shadow, self._shadow_stamp = None, None
        self._ws, self._cfgs = {}, {}
        cfg = self._config(1, 224)
        n = _lib.check(int(_lib.lib.mmu_imgenc_param_count(C.byref(cfg))), "mmu_imgenc_param_count")
        ns = _lib.check(int(_lib.lib.mmu_imgenc_stat_count(C.byref(cfg))), "mmu_imgenc_stat_count")

        def table(fn):
            t = (_lib.ParamEntry * 1024)()
            cnt = _lib.check(fn(C.byref(cfg), t, 1024))
            return [(t[i].name.decode(), int(t[i].offset), int(t[i].numel), int(t[i].rows),
                     int(t[i].cols)) for i in range(cnt)]

        self._table = table(_lib.lib.mmu_imgenc_param_table)
        self._stat_table = table(_lib.lib.mmu_imgenc_stat_table)
        self._flat = torch.zeros(n, dtype=torch.float32)
        self._flat_grad = torch.zeros(n, dtype=torch.float32)
        self._stats = torch.zeros(ns, dtype=torch.float32)
        for name, off, numel, rows, cols in self._table:
            p = nn.Parameter(self._flat[off:off + numel].view(self._shape(name, rows, cols)))
            p._mmu_owner = self
            holder, leaf = _holder_for(self, name)
            holder.register_parameter(leaf, p)
        # every BatchNorm's num_batches_tracked is a 0-dim view of ONE int64 buffer: the forward bumps
        # all of them with a single in-place add instead of 155 tiny launches
        n_bn = sum(1 for name, *_ in self._stat_table if name.endswith("running_var"))
        self._nbt = torch.zeros(n_bn, dtype=torch.long)
        self._nbt_holders = []
        for name, off, numel, rows, cols in self._stat_table:
            holder, leaf = _holder_for(self, name)
            holder.register_buffer(leaf, self._stats[off:off + numel])
            if leaf == "running_var":
                holder.register_buffer("num_batches_tracked", self._nbt[len(self._nbt_holders)])
                self._nbt_holders.append(holder)
        self._rebind(self._flat, self._flat_grad, self._stats)
        self._init_weights()

    @staticmethod
    def _kernel(name):
        return 7 if name == "model.0.weight" else (3 if name.endswith("conv2.weight") else 1)

    def _shape(self, name, rows, cols):
        if cols > 0:
            k = self._kernel(name)
            return (rows, cols // (k * k), k, k)  # OIHW, as nn.Conv2d stores it
        return (rows,)

    def _config(self, B, H):
        cfg = self._cfgs.get((B, H))
        if cfg is None:
            cfg = self._cfgs[(B, H)] = _lib.ImgEncConfig(B, H, (C.c_int * 4)(*self._layers), self._width,
                                                         self._pool[0], self._pool[1], self._pool_max)
        return cfg

    _fresh_shadow = MIMOResNet._fresh_shadow
    invalidate_shadow = MIMOResNet.invalidate_shadow
    _rebind = MIMOResNet._rebind

    def _apply(self, fn, recurse=True):
        flat = fn(self._flat)
        if flat.dtype != torch.float32:
            raise TypeError("the image encoder's master parameters are fp32; choose precision='bf16' "
                            "for the tensor-core path instead of casting the module")
        self._rebind(flat, fn(self._flat_grad), fn(self._stats))
        self._nbt = fn(self._nbt)
        for i, holder in enumerate(self._nbt_holders):
            holder._buffers["num_batches_tracked"] = self._nbt[i]
        return self

    zero_grad = MIMOResNet.zero_grad
    _ensure_grad_views = FlavaFusionTransfomer._ensure_grad_views

    @property
    def _param_list(self):
        return [p for p, _ in self._grad_views]

    def _param_stamp(self):
        return self._flat._version + sum(p._version for p, _ in self._grad_views)

    @torch.no_grad()
    def _init_weights(self):
        """torchvision.models.resnet: kaiming_normal_(fan_out, relu) convolutions, BatchNorm (1, 0)."""
        for name, p in self.named_parameters():
            if p.dim() == 4:
                p.normal_(0.0, math.sqrt(2.0 / (p.shape[0] * p.shape[2] * p.shape[3])))
            elif name.endswith(".weight"):
                p.fill_(1.0)
            else:
                p.zero_()
        for name, off, numel, rows, cols in self._stat_table:
            self._stats[off:off + numel].fill_(1.0 if name.endswith("running_var") else 0.0)

    def _workspace(self, cfg, training):
        key = (cfg.B, cfg.H, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.check(int(_lib.lib.mmu_imgenc_workspace_bytes(C.byref(cfg), int(training))),
                                "mmu_imgenc_workspace_bytes")
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self._flat.device)
        return ws

    def _engine_forward(self, x, training, keep=False):
        if not self._flat.is_cuda:
            raise _lib.MMUError("the image encoder lives on the CPU: call .to('cuda') first -- this "
                                "package has no CPU execution path")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError("expected square images (B, 3, H, H)")
        x = x.to(device=self._flat.device, dtype=torch.float32).contiguous()
        cfg = self._config(x.shape[0], x.shape[2])
        ws = self._workspace(cfg, training)
        tokens = torch.empty(x.shape[0], self._pool[0] * self._pool[1], 2048, device=x.device)
        shadow = self._fresh_shadow()
        bn_training = bool(self.training)
        _lib.check(_lib.lib.mmu_imgenc_forward(C.byref(cfg), self._flat.data_ptr(), _lib.ptr(shadow),
                                               self._stats.data_ptr(), x.data_ptr(), ws.data_ptr(),
                                               ws.numel(), int(bn_training), tokens.data_ptr(),
                                               _lib.stream_ptr()), "mmu_imgenc_forward")
        if bn_training:
            self._nbt += 1
        saved = (cfg, ws, x, shadow, tokens)
        return stamp_workspace(self, ws, saved) if training else saved

    def _engine_backward(self, saved, dtokens):
        cfg, ws, x, shadow, _ = saved
        check_workspace(self, ws, saved)
        self._ensure_grad_views()
        _lib.check(_lib.lib.mmu_imgenc_backward(C.byref(cfg), self._flat.data_ptr(), _lib.ptr(shadow),
                                                self._stats.data_ptr(), x.data_ptr(), ws.data_ptr(),
                                                ws.numel(), dtokens.data_ptr(),
                                                self._flat_grad.data_ptr(), _lib.stream_ptr()),
                   "mmu_imgenc_backward")
        sync = getattr(self, "_flat_sync", None)
        if sync is not None:
            sync.launch(self)

    def forward(self, x):
        anchor = next((p for p, _ in self._grad_views if p.requires_grad), None)
        if self.training and torch.is_grad_enabled() and anchor is not None:
            return _EncoderForward.apply(anchor, self, x)
        # frozen (src/framework.py:281-282) or eval: no backward; BatchNorm follows self.training
        return self._engine_forward(x, training=self.training)[-1]
