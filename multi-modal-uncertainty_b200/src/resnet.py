"""``MIMOResNet`` with the reference's constructor, state-dict keys and call protocol (reference
``src/model.py:17-100``, ``src/layers.py:7-38``), executed by the CUDA ResNet engine
(``csrc/resnet.cu``): the four-view FashionMNIST model of ``train_fashionmnist.py``.

As in the fusion models, every parameter is an individually addressable ``nn.Parameter`` that
views ONE flat fp32 buffer (gradients likewise), and the BatchNorm running statistics are buffer
views of one flat statistics buffer the kernels update in place; reference checkpoints load with
``strict=True``.  ``precision="fp32"`` (default) is the reference's arithmetic;
``precision="bf16"`` runs every convolution on the tcgen05 tensor-core GEMM (bf16 operands, fp32
accumulation, fp32 activations and BatchNorm; 3x3 layers with tap-major columns, the 4-channel
stem with K padded 36 -> 40).  There is no CPU path.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from ._backend import _lib
from .model import FlavaFusionTransfomer, _holder_for, check_workspace, stamp_workspace


class _ResNetForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, model, x):
        ctx.model = model
        ctx.saved = model._engine_forward(x, training=True)
        return ctx.saved[-1]

    @staticmethod
    def backward(ctx, dlogits):
        ctx.model._engine_backward(ctx.saved, dlogits.contiguous())
        return None, None, None


class MIMOResNet(nn.Module):
    """Drop-in for reference ``MIMOResNet(num_channels, emb_dim, out_dim, num_classes)``.
    ``forward(x)`` takes ``(B, E, C, 14, 14)`` (views become channels, src/model.py:83-86) or
    ``(B, C', 14, 14)`` and returns logits ``(B, out_dim, num_classes)``."""

    def __init__(self, num_channels, emb_dim, out_dim, num_classes, precision="fp32"):
        super().__init__()
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' (reference arithmetic) or 'bf16' (tensor cores)")
        self.precision = precision
        self._shadow, self._shadow_stamp = None, None
        self.out_dim, self.num_classes = out_dim, num_classes
        self._cin = num_channels * emb_dim
        self._ws, self._cfgs = {}, {}
        self._last_epi = None
        self.loss = torch.nn.CrossEntropyLoss()  # attribute parity; never called
        cfg = self._config(1)
        n = _lib.check(int(_lib.lib.mmu_resnet_param_count(C.byref(cfg))), "mmu_resnet_param_count")
        ns = _lib.check(int(_lib.lib.mmu_resnet_stat_count(C.byref(cfg))), "mmu_resnet_stat_count")

        def table(fn):
            t = (_lib.ParamEntry * 256)()
            cnt = _lib.check(fn(C.byref(cfg), t, 256))
            return [(t[i].name.decode(), int(t[i].offset), int(t[i].numel), int(t[i].rows),
                     int(t[i].cols)) for i in range(cnt)]

        self._table = table(_lib.lib.mmu_resnet_param_table)
        self._stat_table = table(_lib.lib.mmu_resnet_stat_table)
        self._flat = torch.zeros(n, dtype=torch.float32)
        self._flat_grad = torch.zeros(n, dtype=torch.float32)
        self._stats = torch.zeros(ns, dtype=torch.float32)
        for name, off, numel, rows, cols in self._table:
            p = nn.Parameter(self._flat[off:off + numel].view(self._shape(name, rows, cols)))
            holder, leaf = _holder_for(self, name)
            holder.register_parameter(leaf, p)
        # buffers in the reference's order: running_mean, running_var, num_batches_tracked per BN
        for name, off, numel, rows, cols in self._stat_table:
            holder, leaf = _holder_for(self, name)
            holder.register_buffer(leaf, self._stats[off:off + numel])
            if leaf == "running_var":
                holder.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self._rebind(self._flat, self._flat_grad, self._stats)
        self._init_like_reference()

    def _shape(self, name, rows, cols):
        if name.endswith("conv1.weight") or name.endswith("conv2.weight") or name.endswith("downsample.0.weight"):
            k = 1 if "downsample" in name else 3
            return (rows, cols // (k * k), k, k)  # OIHW, as nn.Conv2d stores it
        return (rows, cols) if cols > 0 else (rows,)

    def _config(self, B):
        cfg = self._cfgs.get(B)
        if cfg is None:
            cfg = self._cfgs[B] = _lib.ResNetConfig(B, self._cin, 14, 14, self.out_dim, self.num_classes)
        return cfg

    # ------------------------------------------------------------------ flat buffers
    def _fresh_shadow(self):
        """bf16 copy of the flat parameters for the tensor-core convolutions, re-cast only when a
        parameter's version counter moved (torch optimisers update the views in place)."""
        if self.precision != "bf16":
            return None
        stamp = self._flat._version + sum(p._version for p, _ in self._grad_views)
        if self._shadow is None or self._shadow.device != self._flat.device:
            self._shadow = torch.empty(self._flat.numel(), dtype=torch.bfloat16, device=self._flat.device)
            self._shadow_stamp = None
        if stamp != self._shadow_stamp:
            _lib.check(_lib.lib.mmu_cast_f32_to_bf16(self._flat.data_ptr(), self._shadow.data_ptr(),
                                                     self._flat.numel(), _lib.stream_ptr()),
                       "mmu_cast_f32_to_bf16")
            self._shadow_stamp = stamp
        return self._shadow

    def invalidate_shadow(self):
        self._shadow_stamp = None

    def _rebind(self, flat, flat_grad, stats):
        self._shadow, self._shadow_stamp = None, None
        self._flat, self._flat_grad, self._stats = flat, flat_grad, stats
        params = dict(self.named_parameters())
        self._grad_views = []
        for name, off, numel, rows, cols in self._table:
            shape = self._shape(name, rows, cols)
            params[name].data = flat[off:off + numel].view(shape)
            params[name].grad = flat_grad[off:off + numel].view(shape)
            self._grad_views.append((params[name], params[name].grad))
        for name, off, numel, rows, cols in self._stat_table:
            holder, leaf = _holder_for(self, name)
            holder._buffers[leaf] = stats[off:off + numel]
        self._ws.clear()

    def _apply(self, fn, recurse=True):
        flat = fn(self._flat)
        if flat.dtype != torch.float32:
            raise TypeError("MIMOResNet runs in fp32 (the reference's configuration)")
        self._rebind(flat, fn(self._flat_grad), fn(self._stats))
        for m in self.modules():
            if "num_batches_tracked" in m._buffers:
                m._buffers["num_batches_tracked"] = fn(m._buffers["num_batches_tracked"])
        return self

    def zero_grad(self, set_to_none: bool = False):
        self._flat_grad.zero_()

    @torch.no_grad()
    def _init_like_reference(self):
        """Seed-for-seed the reference's initial weights: torch's own layers are constructed in
        the reference's order (each consumes the RNG in its default reset_parameters), then
        every conv is re-drawn N(0, sqrt(2 / (k*k*out))) and BN set to (1, 0) in module order
        (src/model.py:33-39)."""
        convs, sd = [], {}

        def conv(name, ci, co, k):
            m = nn.Conv2d(ci, co, k, bias=False)
            convs.append((name, m))

        conv("conv1.weight", self._cin, 64, 3)
        inpl = 64
        for li, (planes, stride) in enumerate(((64, 1), (128, 2)), start=1):
            for bi in range(2):
                pre = f"layer{li}.{bi}."
                # reference order inside _make_layer: the downsample is built BEFORE the block
                if bi == 0 and (stride != 1 or inpl != planes):
                    conv(pre + "downsample.0.weight", inpl, planes, 1)
                conv(pre + "conv1.weight", inpl, planes, 3)
                conv(pre + "conv2.weight", planes, planes, 3)
                inpl = planes
        # the re-initialisation loop walks self.modules(): conv1, layer1.0.conv1, conv2, ...,
        # layer2.0.conv1, conv2, downsample.0, layer2.1.conv1, conv2 (the block registers its
        # downsample after its own convs)
        order = ["conv1.weight", "layer1.0.conv1.weight", "layer1.0.conv2.weight",
                 "layer1.1.conv1.weight", "layer1.1.conv2.weight", "layer2.0.conv1.weight",
                 "layer2.0.conv2.weight", "layer2.0.downsample.0.weight", "layer2.1.conv1.weight",
                 "layer2.1.conv2.weight"]
        by_name = dict(convs)
        for name in order:
            m = by_name[name]
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            m.weight.data.normal_(0, math.sqrt(2.0 / n))
            sd[name] = m.weight
        # MultiHeadFC is created by MIMOResNet.__init__ AFTER ResNet.__init__ returned (:76-78)
        fc = nn.Linear(128, self.num_classes * self.out_dim)
        sd["output_layer.fc.weight"], sd["output_layer.fc.bias"] = fc.weight, fc.bias
        for name, p in self.named_parameters():
            if name in sd:
                p.copy_(sd[name])
            elif name.endswith(".weight"):  # BatchNorm gain
                p.fill_(1.0)
            else:
                p.zero_()
        for name, off, numel, rows, cols in self._stat_table:
            self._stats[off:off + numel].fill_(1.0 if name.endswith("running_var") else 0.0)

    # ------------------------------------------------------------------------ engine
    def _workspace(self, cfg, training):
        key = (cfg.B, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.check(int(_lib.lib.mmu_resnet_workspace_bytes(C.byref(cfg), int(training))),
                                "mmu_resnet_workspace_bytes")
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self._flat.device)
        return ws

    def _engine_forward(self, x, training):
        self._new_forward()
        if not self._flat.is_cuda:
            raise _lib.MMUError("the model lives on the CPU: call .to('cuda') first -- this "
                                "package has no CPU execution path")
        if x.dim() == 5:
            x = x.reshape(x.shape[0], -1, x.shape[3], x.shape[4])
        if x.shape[1] != self._cin or tuple(x.shape[2:]) != (14, 14):
            raise ValueError(f"expected (B, {self._cin}, 14, 14) after folding the views into channels")
        x = x.to(device=self._flat.device, dtype=torch.float32).contiguous()
        cfg = self._config(x.shape[0])
        ws = self._workspace(cfg, training)
        logits = torch.empty(x.shape[0], self.out_dim, self.num_classes, device=x.device)
        shadow = self._fresh_shadow()
        _lib.check(_lib.lib.mmu_resnet_forward(C.byref(cfg), self._flat.data_ptr(), _lib.ptr(shadow),
                                               self._stats.data_ptr(), x.data_ptr(), ws.data_ptr(),
                                               ws.numel(), int(training), logits.data_ptr(),
                                               _lib.stream_ptr()), "mmu_resnet_forward")
        if training:
            for m in self.modules():
                nb = m._buffers.get("num_batches_tracked")
                if nb is not None:
                    nb += 1
        saved = (cfg, ws, x, shadow, logits)
        return stamp_workspace(self, ws, saved) if training else saved

    _ensure_grad_views = FlavaFusionTransfomer._ensure_grad_views
    _labels_key = staticmethod(FlavaFusionTransfomer._labels_key)
    _new_forward = FlavaFusionTransfomer._new_forward

    def _engine_backward(self, saved, dlogits):
        cfg, ws, x, shadow, _ = saved
        check_workspace(self, ws, saved)
        self._ensure_grad_views()
        _lib.check(_lib.lib.mmu_resnet_backward(C.byref(cfg), self._flat.data_ptr(), _lib.ptr(shadow),
                                                self._stats.data_ptr(), x.data_ptr(), ws.data_ptr(),
                                                ws.numel(), dlogits.data_ptr(),
                                                self._flat_grad.data_ptr(), _lib.stream_ptr()),
                   "mmu_resnet_backward")

    # -------------------------------------------------------------- reference protocol
    def forward(self, x):
        if self.training and torch.is_grad_enabled():
            return _ResNetForward.apply(next(self.parameters()), self, x)
        return self._engine_forward(x, training=self.training)[-1]

    def _train_engine(self, x):
        saved = self._engine_forward(x, training=True)
        return saved, saved[-1]

    forward_backward = FlavaFusionTransfomer.forward_backward

    # loss / metric plumbing is the fusion models' (identical semantics, src/model.py:102-112)
    _remember_epilogue = FlavaFusionTransfomer._remember_epilogue
    cached_epilogue = FlavaFusionTransfomer.cached_epilogue
    compute_loss = FlavaFusionTransfomer.compute_loss
