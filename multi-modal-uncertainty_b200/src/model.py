"""Fusion models with the reference's names, constructor signatures, state-dict keys and call
protocol (reference ``src/model.py:225-374``), executed by the sm_100a engine.

``model(x)`` with ``x = (image_features, text_features)`` (either may be ``None`` in eval)
returns logits ``(B, E, C)``; ``model.compute_loss(y_hat, y, eval=False)`` returns the scalar
loss.  All parameters are views into ONE flat fp32 buffer (so the fused AdamW and the gradient
all-reduce see a single contiguous tensor) while staying individually addressable
``nn.Parameter`` objects under the reference's keys, e.g.
``mm_encoder.resblocks.0.attn.in_proj_weight`` -- reference checkpoints load with
``strict=True`` (reference ``src/training_loop.py:72-77``).
"""
import ctypes as C
from typing import Any

import torch
import torch.nn as nn

from ._backend import _lib, ops

# (emb_dim, out_dim) per --model_type of the FashionMNIST scripts (reference src/model.py:8-15)
model_configure = {
    "Vanilla": (4, 1),
    "MIMO-shuffle-instance": (4, 4),
    "MIMO-shuffle-view": (4, 4),
    "MultiHead": (4, 4),
    "MIMO-shuffle-all": (4, 4),
    "single-model-weight-sharing": (1, 1),
}


def __getattr__(name):
    # ``from src.model import MIMOResNet, model_configure, MIMOTransfomer`` (train_fashionmnist.py:17,
    # eval_robustness.py:16): the conv model lives in resnet.py, which imports this module
    if name == "MIMOResNet":
        from .resnet import MIMOResNet
        return MIMOResNet
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

_PREC = {"fp32": _lib.F32, "bf16": _lib.BF16, torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


class _Holder(nn.Module):
    """Parameter container: only exists so that state-dict keys match the reference's tree.
    Containers with numeric children (the reference's ``nn.ModuleList`` / ``nn.Sequential``:
    ``output_layers``, ``resblocks``, ``layer1`` ...) index, iterate and ``len()`` like them."""

    def __getitem__(self, i):
        return self._modules[str(i if i >= 0 else len(self._modules) + i)]

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())


def _holder_for(root, dotted):
    mod = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    return mod, parts[-1]


class _Saved(tuple):
    """What a training-mode engine forward hands to its backward, plus the workspace stamp."""
    stamp = 0


def stamp_workspace(model, ws, saved):
    """The activations a backward needs live in `ws` (ONE workspace per shape): stamp it so that a
    backward whose activations were overwritten by a LATER training-mode forward at the same
    shape fails loudly (``check_workspace``) instead of producing wrong gradients."""
    stamps = model.__dict__.setdefault("_ws_stamp", {})
    out = _Saved(saved)
    out.stamp = stamps[ws.data_ptr()] = stamps.get(ws.data_ptr(), 0) + 1
    return out


def check_workspace(model, ws, saved):
    if model.__dict__.get("_ws_stamp", {}).get(ws.data_ptr()) != getattr(saved, "stamp", None):
        raise _lib.MMUError(
            "backward() of a forward whose saved activations were overwritten: a later "
            "training-mode forward at the same shape ran before this backward (the engine keeps "
            "ONE workspace per shape).  Call backward() before the next training-mode forward, or "
            "run the extra forward after model.eval()")


class _FlavaForward(torch.autograd.Function):
    """One engine call forward, one (staged) engine call backward.  Parameter gradients are
    accumulated by the kernels straight into the model's flat gradient buffer (the ``.grad`` of
    every parameter is a view of it), so backward returns no tensor gradients."""

    @staticmethod
    def forward(ctx, anchor, model, img, txt, idx_img, idx_txt, keep):
        ctx.model = model
        ctx.inputs = model._engine_forward(img, txt, idx_img, idx_txt, keep, training=True)
        return model._logits_train

    @staticmethod
    def backward(ctx, dlogits):
        ctx.model._engine_backward(ctx.inputs, dlogits.contiguous())
        return (None,) * 7


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y, model):
        N, E, _ = y_hat.shape
        dl, _, _, accum = ops.heads_uncertainty_epilogue(
            y_hat, y, 0, grad_scale=1.0 / (N * E), want_grad=True)
        ctx.save_for_backward(dl)
        model._remember_epilogue(y_hat, 0, accum, y)
        return _loss_from_accum(accum)

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def _loss_from_accum(accum):
    o = _lib.ACC_OFF
    loss_sum = accum[o["loss_sum"]:o["loss_sum"] + 1].view(torch.float64)
    n_rows = accum[o["n_rows"]:o["n_rows"] + 1].to(torch.float64)
    return (loss_sum / n_rows).to(torch.float32).reshape(())


class FlavaFusionTransfomer(nn.Module):
    """Drop-in for reference ``FlavaFusionTransfomer`` (src/model.py:225-304).

    Two optional keywords beyond the reference: ``precision`` ("bf16" tensor-core path, default,
    or "fp32" parity path) and ``live_tokens`` (default False = every token position is computed,
    as written).  ``avg_pool`` is REQUIRED, as in the reference (:256).  Device workspaces are
    allocated per (batch, token-count) shape on first use and cached.

    ``live_tokens=True``: without ``avg_pool`` head ``i`` reads token position ``i`` of the
    concatenated sequence (:286-287) and token positions never interact (attention runs over the
    BATCH axis, every other op is row-wise), so positions >= E cannot reach the logits or any
    gradient (~97 % of the as-written step at 197 + 40 tokens).  The
    live-token path gathers only the first E positions of (image ++ text) -- same kernels, same
    rows, bit-identical logits; gradients identical up to the summation order of the split-K
    weight-gradient atomics (the skipped rows contribute exact zeros).  Ignored with ``avg_pool``
    / the CLS variant / ``MIMOTransfomer``, where every position is live.
    """

    _cls_token = False
    live_tokens = False

    def __init__(self,
                 out_dim: int = 1,
                 num_classes: int = 2,
                 image_hidden_size: int = 768,
                 text_hidden_size: int = 768,
                 multimodal_hidden_size: int = 768,
                 multimodal_num_attention_heads: int = 3,
                 multimodal_num_hidden_layers: int = 3,
                 drop: float = 0.0,
                 **kwargs: Any):
        super().__init__()
        self.avg_pool = bool(kwargs["avg_pool"])
        self.out_dim = out_dim
        self.num_classes = num_classes
        self.drop = float(drop)
        self.precision = _PREC[kwargs.get("precision", "bf16")]
        self.live_tokens = bool(kwargs.get("live_tokens", False))
        self._dims = dict(d_img=image_hidden_size, d_txt=text_hidden_size, D=multimodal_hidden_size,
                          n_head=multimodal_num_attention_heads,
                          n_layers=multimodal_num_hidden_layers)
        self._finish_init()

    _group_pool = 0  # MIMOTransfomer: token positions pooled per head

    def _finish_init(self):
        self._ws = {}
        self._live_idx = {}
        self._pack_caps = {}
        self._cfg_cache = {}
        self._last_epi = None
        self._ddp = None
        self._logits_train = None
        self.loss = torch.nn.CrossEntropyLoss()  # kept for attribute parity; never called

        cfg = self._config(1, 1, 1)
        n = int(_lib.lib.mmu_flava_param_count(C.byref(cfg)))
        _lib.check(n, "mmu_flava_param_count")
        table = (_lib.ParamEntry * 512)()
        cnt = _lib.check(_lib.lib.mmu_flava_param_table(C.byref(cfg), table, 512))
        self._table = [(table[i].name.decode(), int(table[i].offset), int(table[i].numel),
                        int(table[i].rows), int(table[i].cols), int(table[i].stage))
                       for i in range(cnt)]
        self._n_stages = int(_lib.lib.mmu_flava_num_stages(C.byref(cfg)))
        self._flat = torch.zeros(n, dtype=torch.float32)
        self._flat_grad = torch.zeros(n, dtype=torch.float32)
        self._register_views()
        self._init_like_reference()

    # ------------------------------------------------------------------ parameters
    def _reference_order(self):
        d = {name: i for i, (name, *_rest) in enumerate(self._table)}
        order = []
        for i in range(self._dims["n_layers"]):
            pre = f"mm_encoder.resblocks.{i}."
            order += [pre + s for s in ("attn.in_proj_weight", "attn.in_proj_bias",
                                        "attn.out_proj.weight", "attn.out_proj.bias",
                                        "ln_1.weight", "ln_1.bias", "mlp.c_fc.weight",
                                        "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias",
                                        "ln_2.weight", "ln_2.bias")]
        order += ["ln_pre.weight", "ln_pre.bias", "ln_post.weight", "ln_post.bias",
                  "image_to_mm_projection.weight", "image_to_mm_projection.bias",
                  "text_to_mm_projection.weight", "text_to_mm_projection.bias"]
        for e in range(self.out_dim):
            order += [f"output_layers.{e}.weight", f"output_layers.{e}.bias"]
        if self._cls_token:
            order.append("class_embeddings")
        assert sorted(order) == sorted(d), "engine parameter table and reference key set differ"
        return [self._table[d[k]] for k in order]

    def _register_views(self):
        # registration order = the reference's parameters() order (optimizer state-dict parity)
        for name, off, numel, rows, cols, _stage in self._reference_order():
            shape = (rows, cols) if cols > 0 else (rows,)
            p = nn.Parameter(self._flat[off:off + numel].view(shape))
            p._mmu_owner = self
            holder, leaf = _holder_for(self, name)
            holder.register_parameter(leaf, p)
        self._rebind(self._flat, self._flat_grad)

    # ---- bf16 shadow of the parameters (tcgen05 GEMM operand), one per model.  It is refreshed
    #      only when a parameter changed: in-place torch ops on any parameter bump that
    #      parameter's version counter, and the fused AdamW (which writes through raw pointers)
    #      updates the shadow itself inside its kernel and re-stamps it.
    def _param_stamp(self):
        return self._flat._version + sum(p._version for p in self._param_list)

    def invalidate_shadow(self):
        """Call after modifying parameters in a way autograd's version counters cannot see
        (writes through ``p.data`` or a raw device pointer)."""
        self._shadow_stamp = None

    def _fresh_shadow(self):
        if self.precision != _PREC["bf16"]:
            return None
        stamp = self._param_stamp()
        if self._shadow is None or self._shadow.device != self._flat.device:
            self._shadow = torch.empty(self._flat.numel(), dtype=torch.bfloat16,
                                       device=self._flat.device)
            self._shadow_stamp = None
        if stamp != self._shadow_stamp:
            _lib.check(_lib.lib.mmu_cast_f32_to_bf16(self._flat.data_ptr(), self._shadow.data_ptr(),
                                                     self._flat.numel(), _lib.stream_ptr()),
                       "mmu_cast_f32_to_bf16")
            self._shadow_stamp = stamp
        return self._shadow

    def _rebind(self, flat, flat_grad):
        self._shadow, self._shadow_stamp = None, None
        self._param_list = list(self.parameters())
        self._flat, self._flat_grad = flat, flat_grad
        params = dict(self.named_parameters())
        self._grad_views = []
        for name, off, numel, rows, cols, _stage in self._table:
            shape = (rows, cols) if cols > 0 else (rows,)
            p = params[name]
            p.data = flat[off:off + numel].view(shape)
            p.grad = flat_grad[off:off + numel].view(shape)
            self._grad_views.append((p, p.grad))
        self._ws.clear()
        self._live_idx = {}

    def train(self, mode: bool = True):
        """``nn.Module.train`` without the recursive ``__setattr__`` walk: a sweep-then-train loop
        toggles eval() / train() every batch, and the generic walk over this model's ~40 parameter
        holders costs 0.3-0.4 ms of host time per toggle pair -- a fifth of a live-token step,
        which is host bound (``tools/host_profile.py``).  ``training`` is a plain instance
        attribute of every module, so writing it directly is equivalent."""
        if not isinstance(mode, bool):
            raise ValueError("training mode is expected to be boolean")
        cache = self.__dict__.get("_mode_modules")
        if cache is None or cache[0] != len(self._modules):
            cache = (len(self._modules), list(self.modules()))
            self.__dict__["_mode_modules"] = cache
        for m in cache[1]:
            m.__dict__["training"] = mode
        return self

    def _ensure_grad_views(self):
        """The kernels accumulate into the flat gradient buffer; ``p.grad`` must be its views.
        ``torch.optim.Optimizer.zero_grad()`` (set_to_none=True, the default) drops them: treat a
        dropped gradient as the zero it stands for and re-attach the view."""
        for p, view in self._grad_views:
            if p.grad is not view:
                if p.grad is None:
                    view.zero_()
                else:
                    view.copy_(p.grad)
                p.grad = view

    def _apply(self, fn, recurse=True):
        flat = fn(self._flat)
        if flat.dtype != torch.float32:
            raise TypeError("the master parameters are fp32; choose precision='bf16' for the "
                            "tensor-core path instead of casting the module")
        grad = fn(self._flat_grad)
        self._rebind(flat, grad)
        return self

    @torch.no_grad()
    def _init_like_reference(self):
        """Same initialisers, called in the same order as the reference constructor would call
        them (torch's own module constructors consume the RNG identically), so a seeded
        construction yields the reference's initial weights."""
        d = self._dims
        sd = {}
        for i in range(d["n_layers"]):
            pre = f"mm_encoder.resblocks.{i}."
            attn = nn.MultiheadAttention(d["D"], d["n_head"])
            sd[pre + "attn.in_proj_weight"] = attn.in_proj_weight
            sd[pre + "attn.in_proj_bias"] = attn.in_proj_bias
            sd[pre + "attn.out_proj.weight"] = attn.out_proj.weight
            sd[pre + "attn.out_proj.bias"] = attn.out_proj.bias
            ln1 = nn.LayerNorm(d["D"])
            fc, proj = nn.Linear(d["D"], 4 * d["D"]), nn.Linear(4 * d["D"], d["D"])
            ln2 = nn.LayerNorm(d["D"])
            for k, m in (("ln_1", ln1), ("mlp.c_fc", fc), ("mlp.c_proj", proj), ("ln_2", ln2)):
                sd[pre + k + ".weight"], sd[pre + k + ".bias"] = m.weight, m.bias
        for k, m in (("ln_pre", nn.LayerNorm(d["D"])), ("ln_post", nn.LayerNorm(d["D"])),
                     ("image_to_mm_projection", nn.Linear(d["d_img"], d["D"])),
                     ("text_to_mm_projection", nn.Linear(d["d_txt"], d["D"]))):
            sd[k + ".weight"], sd[k + ".bias"] = m.weight, m.bias
        for e in range(self.out_dim):
            m = nn.Linear(d["D"], self.num_classes)
            sd[f"output_layers.{e}.weight"], sd[f"output_layers.{e}.bias"] = m.weight, m.bias
        if self._cls_token:
            sd["class_embeddings"] = d["D"] ** -0.5 * torch.randn(d["D"], self.out_dim)
        for name, p in self.named_parameters():
            p.copy_(sd[name])

    def zero_grad(self, set_to_none: bool = False):
        self._flat_grad.zero_()

    # ---------------------------------------------------------------------- engine
    def _config(self, B, l_img, l_txt, max_variants=0):
        key = (B, l_img, l_txt, max_variants, self._group_pool)
        cfg = self._cfg_cache.get(key)
        if cfg is None:
            d = self._dims
            cfg = _lib.FlavaConfig(B, l_img, l_txt, d["d_img"], d["d_txt"], d["D"], d["n_head"],
                                   d["n_layers"], self.out_dim, self.num_classes,
                                   int(self.avg_pool), int(self._cls_token), self.precision,
                                   max_variants, self._group_pool)
            self._cfg_cache[key] = cfg
        return cfg

    def _workspace(self, cfg, training):
        key = (cfg.B, cfg.l_img, cfg.l_txt, cfg.max_variants, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.lib.mmu_flava_workspace_bytes(C.byref(cfg), int(training))
            _lib.check(nbytes, "mmu_flava_workspace_bytes")
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self._flat.device)
            self._ws[key] = ws
        return ws

    @staticmethod
    def _inputs_bf16(img, txt):
        """bf16 host staging: when every given modality arrives as a bf16 tensor it is handed to the
        engine as is (half the host->device bytes).  The bf16 engine rounds its inputs to bf16 in
        the stem anyway, so this is bit-identical to feeding the fp32 values that round to them;
        the fp32 engine widens them exactly.  Mixed or other dtypes are converted to fp32."""
        given = [t for t in (img, txt) if t is not None]
        return bool(given) and all(t.dtype == torch.bfloat16 for t in given)

    def _engine_forward(self, img, txt, idx_img, idx_txt, keep, training):
        if not self._flat.is_cuda:
            raise _lib.MMUError("the model lives on the CPU: call .to('cuda') first -- this "
                                "package has no CPU execution path")
        if not 0.0 <= self.drop < 1.0:
            raise ValueError("drop must be in [0, 1)")
        self._new_forward()
        ref = img if img is not None else txt
        B = ref.shape[0]
        dev = self._flat.device

        src_bf16 = self._inputs_bf16(img, txt)

        def prep(t):
            return None if t is None else t.to(device=dev, dtype=torch.bfloat16 if src_bf16 else torch.float32).contiguous()

        def prep_idx(i):
            return None if i is None else i.to(device=dev, dtype=torch.int32).contiguous()

        img, txt, idx_img, idx_txt = prep(img), prep(txt), prep_idx(idx_img), prep_idx(idx_txt)
        keep = None if keep is None else keep.to(device=dev, dtype=torch.int32).contiguous()
        l_img = img.shape[1] if img is not None else 0
        l_txt = txt.shape[1] if txt is not None else 0
        n_img = (idx_img.numel() if idx_img is not None else l_img) if img is not None else 0
        n_txt = (idx_txt.numel() if idx_txt is not None else l_txt) if txt is not None else 0
        cfg = self._config(B, max(l_img, 1), max(l_txt, 1))
        ws = self._workspace(cfg, training)
        shadow = self._fresh_shadow()
        inp = _lib.FlavaInputs(_lib.ptr(img), _lib.ptr(txt), _lib.ptr(idx_img), _lib.ptr(idx_txt),
                               n_img, n_txt, _lib.ptr(keep), _lib.ptr(shadow))
        inp.src_bf16 = int(src_bf16)
        if training and self.drop > 0.0:
            # nn.Dropout(drop) between c_fc and QuickGELU (reference src/model.py:195-201): the
            # masks are a function of this seed (drawn from torch's CPU generator, so
            # torch.manual_seed reproduces a run) and are regenerated by the backward from `inp`
            inp.drop_p = self.drop
            inp.drop_seed = self.last_dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        logits = torch.empty(B, self.out_dim, self.num_classes, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib.mmu_flava_forward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                              ws.data_ptr(), ws.numel(), int(training),
                                              logits.data_ptr(), _lib.stream_ptr()),
                   "mmu_flava_forward")
        if training:
            self._logits_train = logits
            return stamp_workspace(self, ws, (cfg, inp, ws, (img, txt, idx_img, idx_txt, keep)))  # keeps inputs alive
        return logits

    # ------------------------------------------------------ packed-variant evaluation
    PACK_MAX_POSITIONS = 2048  # token positions per packed pass (bounds the workspace)

    def _variant_segments(self, n_img, n_txt, off_img, off_txt):
        """Rows [begin, end) of the packed sequence feeding each head of one variant: head e
        reads token e of (image tokens ++ text tokens) (src/model.py:286-287) or, with
        ``avg_pool``, the mean over the image / the text tokens (:282-284)."""
        if self.avg_pool:
            return [(off_img, off_img + n_img), (off_txt, off_txt + n_txt)]
        if self.out_dim > n_img + n_txt:
            raise ValueError(f"a variant with {n_img + n_txt} tokens cannot feed {self.out_dim} heads")
        return [((off_img + e, off_img + e + 1) if e < n_img else
                 (off_txt + e - n_img, off_txt + e - n_img + 1)) for e in range(self.out_dim)]

    @torch.no_grad()
    def forward_variants(self, x, variants):
        """Evaluate V token-subset variants of one batch -- ``variants = [(idx_img | None,
        idx_txt | None), ...]``, ``None`` = modality absent, as built by
        ``robustness.robustness_variants`` -- and return logits ``(V, B, E, C)``.

        Equivalent to ``torch.stack([self((img, txt), token_indices=v) for v in variants])``
        (reference loop eval_transformer_robustness.py:99-125), bit for bit, but token positions
        never interact in this model (attention runs over the batch axis; every other op is
        row-wise), so the variants are concatenated along the token axis and evaluated in ONE
        pass: large GEMMs instead of V small ones and ~V times fewer kernel launches."""
        img, txt = x
        if self._cls_token or self.training:
            return torch.stack([self.forward((img if ii is not None else None,
                                              txt if it is not None else None),
                                             token_indices=(ii, it)) for ii, it in variants])
        if self._skips_dead_tokens():
            live = []
            for ii, it in variants:
                _, _, li, lt = self._live_subset(img if ii is not None else None,
                                                 txt if it is not None else None, ii, it)
                live.append((li, lt))
            variants = live
        out, chunk, pos = [], [], 0
        for v in variants:
            n = (len(v[0]) if v[0] is not None else 0) + (len(v[1]) if v[1] is not None else 0)
            if chunk and pos + n > self.PACK_MAX_POSITIONS:
                out.append(self._forward_packed(img, txt, chunk))
                chunk, pos = [], 0
            chunk.append(v)
            pos += n
        out.append(self._forward_packed(img, txt, chunk))
        return out[0] if len(out) == 1 else torch.cat(out)

    def _forward_packed(self, img, txt, variants):
        self._new_forward()
        if not self._flat.is_cuda:
            raise _lib.MMUError("the model lives on the CPU: call .to('cuda') first -- this "
                                "package has no CPU execution path")
        dev = self._flat.device
        V = len(variants)
        ni = [len(v[0]) if v[0] is not None else 0 for v in variants]
        nt = [len(v[1]) if v[1] is not None else 0 for v in variants]
        Ni, Nt = sum(ni), sum(nt)
        if (Ni and img is None) or (Nt and txt is None) or Ni + Nt == 0:
            raise ValueError("variants index a modality that was not given")
        segs, oi, ot = [], 0, Ni
        for a, b in zip(ni, nt):
            segs += self._variant_segments(a, b, oi, ot)
            oi, ot = oi + a, ot + b

        def cat_idx(k, total):
            if total == 0:
                return None
            return torch.cat([v[k].to(torch.int32) for v in variants if v[k] is not None and len(v[k])])

        # one small host tensor -> one H2D copy: [segments | idx_img | idx_txt]
        seg_t = torch.tensor(segs, dtype=torch.int32).reshape(-1)
        parts = [seg_t] + [t for t in (cat_idx(0, Ni), cat_idx(1, Nt)) if t is not None]
        # pinned staging (torch's caching host allocator recycles it behind a stream event): a
        # pageable source would make cudaMemcpyAsync synchronise the stream first
        staged = torch.empty(sum(t.numel() for t in parts), dtype=torch.int32, pin_memory=True)
        torch.cat(parts, out=staged)
        packed = staged.to(dev, non_blocking=True)
        seg_d = packed[: seg_t.numel()]
        idx_img = packed[seg_t.numel(): seg_t.numel() + Ni] if Ni else None
        idx_txt = packed[seg_t.numel() + Ni:] if Nt else None

        src_bf16 = self._inputs_bf16(img if Ni else None, txt if Nt else None)

        def prep(t):
            return None if t is None else t.to(device=dev, dtype=torch.bfloat16 if src_bf16 else torch.float32).contiguous()

        img, txt = prep(img) if Ni else None, prep(txt) if Nt else None
        B = (img if img is not None else txt).shape[0]
        # grow-only capacities so that every sweep of this batch size shares one workspace
        cap = self._pack_caps.get(B, (1, 1, 2))
        cap = (max(cap[0], Ni, 1), max(cap[1], Nt, 1), max(cap[2], V))
        self._pack_caps[B] = cap
        for key in [k for k in self._ws if k[0] == B and k[3] > 1 and k[1:4] != cap]:
            del self._ws[key]  # superseded by the larger workspace
        cfg = self._config(B, cap[0], cap[1], cap[2])
        ws = self._workspace(cfg, False)
        inp = _lib.FlavaInputs(_lib.ptr(img), _lib.ptr(txt), _lib.ptr(idx_img), _lib.ptr(idx_txt),
                               Ni, Nt, 0, _lib.ptr(self._fresh_shadow()),
                               img.shape[1] if img is not None else 0,
                               txt.shape[1] if txt is not None else 0, V, _lib.ptr(seg_d))
        inp.src_bf16 = int(src_bf16)
        logits = torch.empty(V, B, self.out_dim, self.num_classes, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib.mmu_flava_forward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                              ws.data_ptr(), ws.numel(), 0, logits.data_ptr(),
                                              _lib.stream_ptr()), "mmu_flava_forward")
        return logits

    def _engine_backward(self, saved, dlogits):
        cfg, inp, ws, _alive = saved
        check_workspace(self, ws, saved)
        self._ensure_grad_views()
        if self._ddp is not None:
            self._ddp.backward(self, cfg, inp, ws, dlogits)
            return
        _lib.check(_lib.lib.mmu_flava_backward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                               ws.data_ptr(), ws.numel(), dlogits.data_ptr(),
                                               self._flat_grad.data_ptr(), 0, self._n_stages,
                                               _lib.stream_ptr()), "mmu_flava_backward")

    def backward_stages(self, cfg, inp, ws, dlogits, begin, end):
        _lib.check(_lib.lib.mmu_flava_backward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                               ws.data_ptr(), ws.numel(), dlogits.data_ptr(),
                                               self._flat_grad.data_ptr(), begin, end,
                                               _lib.stream_ptr()), "mmu_flava_backward")

    def stage_ranges(self):
        """[(begin, end)] element ranges of the flat gradient finished by each backward stage."""
        out = []
        for st in range(self._n_stages):
            offs = [(o, o + n) for _nm, o, n, _r, _c, s in self._table if s == st]
            out.append((min(a for a, _ in offs), max(b for _, b in offs)))
        return out

    # ------------------------------------------------------------------ live-token path
    def _skips_dead_tokens(self):
        return self.live_tokens and not self.avg_pool and not self._cls_token and self._group_pool == 0

    def _live_subset(self, img, txt, idx_img, idx_txt):
        """(img, txt, idx_img, idx_txt) restricted to the first E positions of (image ++ text):
        the only ones head i = 0..E-1 reads (src/model.py:286-287)."""
        E = self.out_dim
        n_img = (len(idx_img) if idx_img is not None else img.shape[1]) if img is not None else 0
        n_txt = (len(idx_txt) if idx_txt is not None else txt.shape[1]) if txt is not None else 0
        if n_img + n_txt < E:
            raise ValueError(f"{n_img + n_txt} tokens cannot feed {E} heads")
        k_img = min(E, n_img)
        k_txt = E - k_img
        dev = self._flat.device

        def head(idx, k):
            if idx is not None:
                return idx[:k]
            t = self._live_idx.get(k)
            if t is None:
                t = self._live_idx[k] = torch.arange(k, dtype=torch.int32, device=dev)
            return t
        return (img if k_img else None, txt if k_txt else None,
                head(idx_img, k_img) if k_img else None, head(idx_txt, k_txt) if k_txt else None)

    # -------------------------------------------------------------- reference protocol
    def forward(self, x, token_indices=None, keep_mask=None):
        """``x = (image_features, text_features)``.  ``token_indices = (idx_img, idx_txt)`` runs
        the model on token subsets gathered on device (robustness sweeps); ``keep_mask`` is an
        int (B, 2) modality keep mask (0 zero-fills that modality for that sample)."""
        img, txt = x
        idx_img, idx_txt = token_indices if token_indices is not None else (None, None)
        if self._skips_dead_tokens():
            img, txt, idx_img, idx_txt = self._live_subset(img, txt, idx_img, idx_txt)
        if self.training and torch.is_grad_enabled():
            anchor = next(self.parameters())
            return _FlavaForward.apply(anchor, self, img, txt, idx_img, idx_txt, keep_mask)
        return self._engine_forward(img, txt, idx_img, idx_txt, keep_mask, training=False)

    def _train_engine(self, x):
        img, txt = x
        idx_img = idx_txt = None
        if self._skips_dead_tokens():
            img, txt, idx_img, idx_txt = self._live_subset(img, txt, None, None)
        saved = self._engine_forward(img, txt, idx_img, idx_txt, None, training=True)
        return saved, self._logits_train

    @torch.no_grad()
    def forward_backward(self, x, y):
        """``y_hat = model(x); loss = model.compute_loss(y_hat, y); loss.backward()`` of a training
        step in one call that bypasses the autograd engine: engine forward, fused CE + gradient
        epilogue, engine backward into the flat gradient buffer -- the same kernels in the same
        order.  This is the form a CUDA graph can capture (``graphs.GraphedTrainStep``): the
        autograd engine's cross-stream bookkeeping creates dependencies on uncaptured work.
        Returns (logits, loss)."""
        saved, logits = self._train_engine(x)
        y2 = y.reshape(logits.shape[0], -1)
        if y2.shape[1] not in (1, logits.shape[1]):
            raise ValueError("labels must be (B,) or (B, E)")
        N, E, _ = logits.shape
        dl, _, _, accum = ops.heads_uncertainty_epilogue(logits, y2.contiguous(), 0,
                                                         grad_scale=1.0 / (N * E), want_grad=True)
        self._remember_epilogue(logits, 0, accum, y2)
        self._engine_backward(saved, dl)
        logits._mmu_owner = self   # lets metrics.acc reuse this epilogue's accumulator
        return logits, _loss_from_accum(accum)

    # The loss epilogue's accumulator is reused by ``metrics.acc`` for the SAME logits and labels.
    # The engine writes logits through raw pointers (version counters do not move) and the caching
    # allocator recycles addresses, so the cache entry is also tied to a forward GENERATION: every
    # engine forward bumps ``_fwd_gen`` and drops the entry.
    @staticmethod
    def _labels_key(labels):
        return None if labels is None else (labels.data_ptr(), labels.numel(), labels._version)

    def _remember_epilogue(self, y_hat, mode, accum, labels=None):
        self._last_epi = (y_hat.data_ptr(), y_hat._version, tuple(y_hat.shape), mode,
                          getattr(self, "_fwd_gen", 0), self._labels_key(labels), accum)

    def cached_epilogue(self, y_hat, mode, labels=None):
        e = self._last_epi
        if e is not None and e[0] == y_hat.data_ptr() and e[1] == y_hat._version \
                and e[2] == tuple(y_hat.shape) and e[3] == mode and e[4] == getattr(self, "_fwd_gen", 0) \
                and e[5] == self._labels_key(labels):
            return e[6]
        return None

    def _new_forward(self):
        self._fwd_gen = getattr(self, "_fwd_gen", 0) + 1
        self._last_epi = None

    def compute_loss(self, y_hat, y, eval=False):
        """Reference src/model.py:293-304: train -> mean CE over the (B*E) head rows against the
        tiled labels; eval -> CE on the head-mean logits."""
        assert y.shape[0] == y_hat.shape[0]
        y_hat = y_hat if y_hat.is_contiguous() else y_hat.contiguous()
        if not eval:
            y2 = y.reshape(y_hat.shape[0], -1)
            if y2.shape[1] not in (1, y_hat.shape[1]):
                raise ValueError("labels must be (B,) or (B, E)")
            if y_hat.requires_grad:
                return _LossFn.apply(y_hat, y2.contiguous(), self)
            _, _, _, accum = ops.heads_uncertainty_epilogue(y_hat, y2.contiguous(), 0)
            self._remember_epilogue(y_hat, 0, accum, y2)
            return _loss_from_accum(accum)
        _, _, _, accum = ops.heads_uncertainty_epilogue(y_hat.detach(), y.reshape(-1).contiguous(), 1)
        self._remember_epilogue(y_hat, 1, accum, y)
        return _loss_from_accum(accum)


class FlavaFusionTransfomerwithCLSToken(FlavaFusionTransfomer):
    """Drop-in for reference ``FlavaFusionTransfomerwithCLSToken`` (src/model.py:306-374): E
    learned class rows are prepended to every sample's sequence and head i reads row i."""

    _cls_token = True

    def __init__(self,
                 out_dim: int = 1,
                 num_classes: int = 2,
                 image_hidden_size: int = 768,
                 text_hidden_size: int = 768,
                 multimodal_hidden_size: int = 768,
                 multimodal_num_attention_heads: int = 3,
                 multimodal_num_hidden_layers: int = 3,
                 drop: float = 0.1,
                 **kwargs: Any):
        super().__init__(out_dim, num_classes, image_hidden_size, text_hidden_size,
                         multimodal_hidden_size, multimodal_num_attention_heads,
                         multimodal_num_hidden_layers, drop, **kwargs)


class MIMOTransfomer(FlavaFusionTransfomer):
    """Drop-in for reference ``MIMOTransfomer`` (src/model.py:114-171): the four-view FashionMNIST
    transformer.  ``x`` is ``(B, E, C, H, W)``; every (view, channel) image is one token of
    ``H*W`` pixels, projected to ``hidden_size``, run through the same batch-axis-attention blocks
    as the FLAVA fusion model, and head ``e`` reads the mean over the ``C`` tokens of view ``e``.
    It runs on the same engine (single modality, ``group_pool`` head wiring); the 196-pixel
    projection is not 16-byte-row aligned in bf16, so the engine keeps that small stem GEMM in fp32.
    """

    def __init__(self,
                 out_dim,
                 num_classes,
                 hidden_size,
                 image_dim=14 * 14,
                 multimodal_num_hidden_layers=3,
                 multimodal_num_attention_heads=3,
                 drop=0,
                 **kwargs: Any):
        nn.Module.__init__(self)
        self.avg_pool = False
        self.out_dim = out_dim
        self.num_classes = num_classes
        self.drop = float(drop)
        self.precision = _PREC[kwargs.get("precision", "bf16")]
        self._dims = dict(d_img=image_dim, d_txt=0, D=hidden_size,
                          n_head=multimodal_num_attention_heads,
                          n_layers=multimodal_num_hidden_layers)
        self._group_pool = 1
        self._finish_init()

    def _block_keys(self, i):
        pre = f"mm_encoder.resblocks.{i}."
        return [pre + s for s in ("attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                                  "attn.out_proj.bias", "ln_1.weight", "ln_1.bias",
                                  "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight",
                                  "mlp.c_proj.bias", "ln_2.weight", "ln_2.bias")]

    def _reference_order(self):
        d = {name: i for i, (name, *_rest) in enumerate(self._table)}
        order = ["image_to_mm_projection.weight", "image_to_mm_projection.bias"]
        for i in range(self._dims["n_layers"]):
            order += self._block_keys(i)
        for e in range(self.out_dim):
            order += [f"output_layers.{e}.weight", f"output_layers.{e}.bias"]
        order += ["ln_pre.weight", "ln_pre.bias", "ln_post.weight", "ln_post.bias"]
        assert sorted(order) == sorted(d), "engine parameter table and reference key set differ"
        return [self._table[d[k]] for k in order]

    @torch.no_grad()
    def _init_like_reference(self):
        """Initialisers in the reference constructor's order (src/model.py:126-137): projection,
        blocks, heads; the two LayerNorms consume no random numbers."""
        d = self._dims
        sd = {}
        m = nn.Linear(d["d_img"], d["D"])
        sd["image_to_mm_projection.weight"], sd["image_to_mm_projection.bias"] = m.weight, m.bias
        for i in range(d["n_layers"]):
            pre = f"mm_encoder.resblocks.{i}."
            attn = nn.MultiheadAttention(d["D"], d["n_head"])
            sd[pre + "attn.in_proj_weight"] = attn.in_proj_weight
            sd[pre + "attn.in_proj_bias"] = attn.in_proj_bias
            sd[pre + "attn.out_proj.weight"] = attn.out_proj.weight
            sd[pre + "attn.out_proj.bias"] = attn.out_proj.bias
            ln1 = nn.LayerNorm(d["D"])
            fc, proj = nn.Linear(d["D"], 4 * d["D"]), nn.Linear(4 * d["D"], d["D"])
            ln2 = nn.LayerNorm(d["D"])
            for k, mod in (("ln_1", ln1), ("mlp.c_fc", fc), ("mlp.c_proj", proj), ("ln_2", ln2)):
                sd[pre + k + ".weight"], sd[pre + k + ".bias"] = mod.weight, mod.bias
        for e in range(self.out_dim):
            mod = nn.Linear(d["D"], self.num_classes)
            sd[f"output_layers.{e}.weight"], sd[f"output_layers.{e}.bias"] = mod.weight, mod.bias
        for k in ("ln_pre", "ln_post"):
            mod = nn.LayerNorm(d["D"])
            sd[k + ".weight"], sd[k + ".bias"] = mod.weight, mod.bias
        for name, p in self.named_parameters():
            p.copy_(sd[name])

    def forward(self, x, keep_mask=None):
        b, e, c, h, w = x.shape
        if e != self.out_dim or h * w != self._dims["d_img"]:
            raise ValueError(f"expected (B, {self.out_dim}, C, H, W) with H*W = {self._dims['d_img']}")
        self._group_pool = c
        tokens = x.reshape(b, e * c, h * w)
        return super().forward((tokens, None), keep_mask=keep_mask)

    def _train_engine(self, x):
        b, e, c, h, w = x.shape
        if e != self.out_dim or h * w != self._dims["d_img"]:
            raise ValueError(f"expected (B, {self.out_dim}, C, H, W) with H*W = {self._dims['d_img']}")
        self._group_pool = c
        return super()._train_engine((x.reshape(b, e * c, h * w), None))

    def forward_variants(self, x, variants):
        raise NotImplementedError("token-subset variants are a FLAVA-fusion sweep; the FashionMNIST "
                                  "sweep zero-fills views (robustness.run_view_robustness)")
