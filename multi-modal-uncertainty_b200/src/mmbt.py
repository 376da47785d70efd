"""``MultimodalBertClf`` with the reference's constructor, state-dict keys and call protocol
(reference ``src/mmbt.py:238-262``; encoder ``:86-234``; ``ImageBertEmbeddings`` ``:47-84``),
executed by the CUDA MMBT engine (``csrc/mmbt.cu``).

* ``model(txt, mask, segment, img)``, ``forward_img_only``, ``forward_txt_only`` and
  ``forward_control(txt, mask, segment, img, control_modal)`` are ONE engine path: each is an
  index list over the full sequence ``[CLS] img.. [SEP] | text..`` (``forward_indices``).
* ``img`` may be raw images ``(B, 3, H, W)`` -- they go through ``enc.img_encoder`` -- or already
  pooled image tokens ``(B, num_image_embeds, img_hidden_sz)`` (the cached tokens of a frozen image
  encoder).
* Every parameter is an ``nn.Parameter`` view of one flat fp32 buffer (gradients likewise); the
  tensors ``ImageBertEmbeddings`` shares with the text embeddings (``src/mmbt.py:51-55``) appear
  under both prefixes in ``state_dict()`` exactly as in the reference, so reference checkpoints
  load with ``strict=True``.
* The BERT arithmetic follows the published definitions of the reference's un-vendored dependency
  ``pytorch_pretrained_bert`` (restated for the tests in ``oracle/bert_restated.py``).
* Dropout in ``train()`` mode as in the reference: BERT's hidden / attention-probability dropouts
  (``BertConfig`` defaults 0.1; ``bert_config`` keys ``hidden_dropout_prob`` /
  ``attention_probs_dropout_prob`` or ``args.bert_dropout`` change them) and
  ``ImageBertEmbeddings.dropout`` (``args.dropout``, ``src/mmbt.py:56,82``).  Masks come from the
  engine's counter-based generator (``csrc/dropout.cuh``; statistical parity with torch's Philox
  draws, bit-exact against ``oracle/dropout.py``) and are regenerated in the backward, also
  inside the fused attention kernels (bf16, head_dim 64, S <= 512).
There is no CPU path.
"""
import ctypes as C

import torch
import torch.nn as nn

from ._backend import _lib, ops
from .model import (FlavaFusionTransfomer, _Holder, _holder_for, _loss_from_accum, check_workspace,
                    stamp_workspace)

_PREC = {"fp32": 0, "bf16": 1}

#: architecture of the checkpoints ``BertModel.from_pretrained(name)`` would fetch (no network here:
#: weights are random-initialised exactly like ``init_bert_weights`` and meant to be overwritten by
#: ``load_state_dict``).  ``args.bert_config`` (a dict with the same keys) overrides the lookup.
# hidden_dropout_prob / attention_probs_dropout_prob: 0.1 in the pretrained checkpoints' configs
# (and BertConfig's defaults), i.e. what the reference's BertModel applies in train() mode
BERT_CONFIGS = {
    "bert-base-uncased": dict(vocab=30522, D=768, n_head=12, n_layers=12, d_ff=3072, max_pos=512,
                              n_types=2, init_range=0.02, hidden_dropout_prob=0.1,
                              attention_probs_dropout_prob=0.1),
    "bert-large-uncased": dict(vocab=30522, D=1024, n_head=16, n_layers=24, d_ff=4096, max_pos=512,
                               n_types=2, init_range=0.02, hidden_dropout_prob=0.1,
                               attention_probs_dropout_prob=0.1),
}


class _MmbtForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, model, txt, mask, segment, indices):
        ctx.model = model
        ctx.saved = model._engine_forward(tokens, txt, mask, segment, indices, training=True)
        ctx.need_dimg = tokens.requires_grad
        return ctx.saved[-1]

    @staticmethod
    def backward(ctx, dlogits):
        dimg = ctx.model._engine_backward(ctx.saved, dlogits.contiguous(), ctx.need_dimg)
        return dimg, None, None, None, None, None


class _CELoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y, model):
        N = y_hat.shape[0]
        dl, _, _, accum = ops.heads_uncertainty_epilogue(y_hat.view(N, 1, -1), y.view(N, 1), 0,
                                                         grad_scale=1.0 / N, want_grad=True)
        ctx.save_for_backward(dl.view_as(y_hat))
        model._remember_epilogue(y_hat, 0, accum)
        return _loss_from_accum(accum)

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


class MultimodalBertClf(nn.Module):
    """Drop-in for reference ``MultimodalBertClf(args)``.  ``args`` needs what the reference reads:
    ``bert_model``, ``hidden_sz``, ``img_hidden_sz``, ``num_image_embeds``, ``img_embed_pool_type``,
    ``dropout``, ``n_classes``, ``vocab.stoi`` -- plus, optionally, ``precision`` ("bf16" default,
    "fp32" parity path), ``bert_config`` and ``img_encoder`` ("native", the default: the CUDA
    ResNet-152 ``ImageEncoder``; ``None``: a tokens-only model fed pooled ``(B, N, img_hidden_sz)``
    image tokens)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        bc = dict(getattr(args, "bert_config", None) or BERT_CONFIGS[args.bert_model])
        if bc["D"] != args.hidden_sz:
            raise ValueError("args.hidden_sz must equal the BERT hidden size")
        self.precision = _PREC[getattr(args, "precision", "bf16")]
        self._bc = bc
        # Dropout of a training-mode forward (reference: BertModel's hidden / attention-probability
        # dropouts, BertConfig defaults 0.1, and ImageBertEmbeddings.dropout = args.dropout,
        # src/mmbt.py:56).  ``args.bert_dropout`` (optional) overrides both BERT probabilities.
        bd = getattr(args, "bert_dropout", None)
        self.drop_hidden = float(bc.get("hidden_dropout_prob", 0.1) if bd is None else bd)
        self.drop_attn = float(bc.get("attention_probs_dropout_prob", 0.1) if bd is None else bd)
        self.drop_img = float(getattr(args, "dropout", 0.0) or 0.0)
        for pr in (self.drop_hidden, self.drop_attn, self.drop_img):
            if not 0.0 <= pr < 1.0:
                raise ValueError("dropout probabilities must be in [0, 1)")
        self.last_dropout_seed = None
        self._n_img = int(args.num_image_embeds)
        self._d_img = int(args.img_hidden_sz)
        self._cls_id = int(args.vocab.stoi["[CLS]"])
        self._sep_id = int(args.vocab.stoi["[SEP]"])
        self.loss = nn.CrossEntropyLoss()  # attribute parity; never called
        self._ws, self._cfgs, self._idx_cache = {}, {}, {}
        self._last_epi = None
        self._shadow, self._shadow_stamp = None, None

        cfg = self._config(1, 1)
        n = _lib.check(int(_lib.lib.mmu_mmbt_param_count(C.byref(cfg))), "mmu_mmbt_param_count")
        t = (_lib.ParamEntry * 1024)()
        cnt = _lib.check(_lib.lib.mmu_mmbt_param_table(C.byref(cfg), t, 1024), "mmu_mmbt_param_table")
        self._table = {t[i].name.decode(): (int(t[i].offset), int(t[i].numel), int(t[i].rows), int(t[i].cols))
                       for i in range(cnt)}
        self._flat = torch.zeros(n, dtype=torch.float32)
        self._flat_grad = torch.zeros(n, dtype=torch.float32)
        # sub-module skeleton in the reference's registration order (src/mmbt.py:90-96, :241-242)
        enc = _Holder()
        self.add_module("enc", enc)
        for sub in ("txt_embeddings", "img_embeddings", "img_encoder", "encoder", "pooler"):
            if sub == "img_encoder":
                ie = getattr(args, "img_encoder", "native")
                if isinstance(ie, str) and ie == "native":
                    from .image_encoder import ImageEncoder
                    ie = ImageEncoder(args)
                enc.add_module(sub, ie)  # None: tokens-only model (no enc.img_encoder.* keys)
            else:
                enc.add_module(sub, _Holder())
        for name in self._reference_order():
            off, numel, rows, cols = self._table[name]
            p = nn.Parameter(self._flat[off:off + numel].view(self._shape(rows, cols)))
            p._mmu_owner = self
            holder, leaf = _holder_for(self, name)
            holder.register_parameter(leaf, p)
        # ImageBertEmbeddings shares these modules with the text embeddings (src/mmbt.py:51-55)
        for shared in ("position_embeddings", "token_type_embeddings", "word_embeddings", "LayerNorm"):
            enc.img_embeddings.add_module(shared, enc.txt_embeddings._modules[shared])
        self._rebind(self._flat, self._flat_grad)
        self._init_like_reference()

    # ------------------------------------------------------------------ layout helpers
    def _reference_order(self):
        names = ["enc.txt_embeddings.word_embeddings.weight", "enc.txt_embeddings.position_embeddings.weight",
                 "enc.txt_embeddings.token_type_embeddings.weight", "enc.txt_embeddings.LayerNorm.weight",
                 "enc.txt_embeddings.LayerNorm.bias", "enc.img_embeddings.img_embeddings.weight",
                 "enc.img_embeddings.img_embeddings.bias"]
        for i in range(self._bc["n_layers"]):
            pre = f"enc.encoder.layer.{i}."
            for mod in ("attention.self.query", "attention.self.key", "attention.self.value",
                        "attention.output.dense", "attention.output.LayerNorm", "intermediate.dense",
                        "output.dense", "output.LayerNorm"):
                names += [pre + mod + ".weight", pre + mod + ".bias"]
        names += ["enc.pooler.dense.weight", "enc.pooler.dense.bias", "clf.weight", "clf.bias"]
        assert sorted(names) == sorted(self._table)
        return names

    @staticmethod
    def _shape(rows, cols):
        return (rows, cols) if cols > 0 else (rows,)

    def _config(self, B, S_txt, max_seq=0):
        key = (B, S_txt, max_seq)
        cfg = self._cfgs.get(key)
        if cfg is None:
            bc = self._bc
            cfg = self._cfgs[key] = _lib.MmbtConfig(
                B, S_txt, self._n_img, self._d_img, bc["D"], bc["n_head"], bc["n_layers"], bc["d_ff"],
                bc["vocab"], bc["max_pos"], bc["n_types"], int(self.args.n_classes), self._cls_id,
                self._sep_id, self.precision, max_seq, self.drop_hidden, self.drop_attn, self.drop_img, 0)
        return cfg

    def _own_parameters(self):
        """(name, parameter) of the tensors in this module's flat buffer (not the image encoder's)."""
        return [(k, p) for k, p in self.named_parameters() if not k.startswith("enc.img_encoder.")]

    def _rebind(self, flat, flat_grad):
        self._shadow, self._shadow_stamp = None, None
        self._flat, self._flat_grad = flat, flat_grad
        params = dict(self._own_parameters())
        self._grad_views = []
        for name, (off, numel, rows, cols) in self._table.items():
            shape = self._shape(rows, cols)
            params[name].data = flat[off:off + numel].view(shape)
            params[name].grad = flat_grad[off:off + numel].view(shape)
            self._grad_views.append((params[name], params[name].grad))
        self._param_list = [p for p, _ in self._grad_views]
        self._ws.clear()
        self._idx_cache.clear()

    def _apply(self, fn, recurse=True):
        flat = fn(self._flat)
        if flat.dtype != torch.float32:
            raise TypeError("the master parameters are fp32; choose precision='bf16' for the "
                            "tensor-core path instead of casting the module")
        self._rebind(flat, fn(self._flat_grad))
        ie = self.enc._modules["img_encoder"]
        if ie is not None:
            ie._apply(fn)
        return self

    def zero_grad(self, set_to_none: bool = False):
        self._flat_grad.zero_()
        ie = self.enc._modules["img_encoder"]
        if ie is not None:
            ie.zero_grad(set_to_none)

    _ensure_grad_views = FlavaFusionTransfomer._ensure_grad_views
    _param_stamp = FlavaFusionTransfomer._param_stamp
    invalidate_shadow = FlavaFusionTransfomer.invalidate_shadow
    _fresh_shadow = FlavaFusionTransfomer._fresh_shadow
    _remember_epilogue = FlavaFusionTransfomer._remember_epilogue
    cached_epilogue = FlavaFusionTransfomer.cached_epilogue
    _labels_key = staticmethod(FlavaFusionTransfomer._labels_key)
    _new_forward = FlavaFusionTransfomer._new_forward

    @torch.no_grad()
    def _init_like_reference(self):
        """pytorch_pretrained_bert's ``init_bert_weights`` for the BERT tensors (N(0, range) weights,
        LayerNorm (1, 0), zero biases), ``nn.Linear`` defaults for ``img_embeddings`` and ``clf``
        (src/mmbt.py:50, :242)."""
        r = self._bc["init_range"]
        for name, p in self._own_parameters():
            if name.startswith("enc.img_embeddings.img_embeddings") or name.startswith("clf"):
                continue
            if "LayerNorm.weight" in name:
                p.fill_(1.0)
            elif p.dim() == 1:
                p.zero_()
            else:
                p.normal_(0.0, r)
        for w, b in ((self.enc.img_embeddings.img_embeddings.weight, self.enc.img_embeddings.img_embeddings.bias),
                     (self.clf.weight, self.clf.bias)):
            lin = nn.Linear(w.shape[1], w.shape[0])
            w.copy_(lin.weight)
            b.copy_(lin.bias)

    # ------------------------------------------------------------------------ engine
    def _workspace(self, cfg, training):
        key = (cfg.B, cfg.S_txt, cfg.max_seq, bool(training))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.check(int(_lib.lib.mmu_mmbt_workspace_bytes(C.byref(cfg), int(training))),
                                "mmu_mmbt_workspace_bytes")
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self._flat.device)
        return ws

    def _device_indices(self, indices):
        if indices is None:
            return None
        if torch.is_tensor(indices) and indices.dim() == 2:  # one list per sample
            return indices.to(device=self._flat.device, dtype=torch.int32).contiguous()
        key = tuple(int(i) for i in indices)
        t = self._idx_cache.get(key)
        if t is None:
            if len(self._idx_cache) > 256:
                self._idx_cache.clear()
            t = self._idx_cache[key] = torch.tensor(key, dtype=torch.int32, device=self._flat.device)
        return t

    def _engine_forward(self, tokens, txt, mask, segment, indices, training):
        self._new_forward()
        if not self._flat.is_cuda:
            raise _lib.MMUError("the model lives on the CPU: call .to('cuda') first -- this "
                                "package has no CPU execution path")
        dev = self._flat.device
        B, S_txt = txt.shape
        if tuple(tokens.shape) != (B, self._n_img, self._d_img):
            raise ValueError(f"image tokens must be (B, {self._n_img}, {self._d_img})")
        tokens = tokens.detach().to(device=dev, dtype=torch.float32).contiguous()
        txt, mask, segment = (t.to(device=dev, dtype=torch.int64).contiguous() for t in (txt, mask, segment))
        idx = self._device_indices(indices)
        # per-sample lists (packed short variants): the workspace is sized to the lists, not to the
        # full sequence of the replicated batch
        cfg = self._config(B, S_txt, int(idx.shape[-1]) if (idx is not None and idx.dim() == 2) else 0)
        ws = self._workspace(cfg, training)
        shadow = self._fresh_shadow()
        per_sample = int(idx is not None and idx.dim() == 2)
        if per_sample and idx.shape[0] != B:
            raise ValueError("per-sample index lists must be (B, S)")
        seed = 0
        if training and (self.drop_hidden > 0 or self.drop_attn > 0 or self.drop_img > 0):
            # masks are a function of this seed (torch's CPU generator: torch.manual_seed reproduces
            # a run); the backward regenerates them from the same value
            seed = self.last_dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        inp = _lib.MmbtInputs(txt.data_ptr(), mask.data_ptr(), segment.data_ptr(), tokens.data_ptr(),
                              _lib.ptr(idx), 0 if idx is None else idx.shape[-1], per_sample,
                              _lib.ptr(shadow), None, seed)
        logits = torch.empty(B, int(self.args.n_classes), device=dev, dtype=torch.float32)
        _lib.check(_lib.lib.mmu_mmbt_forward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                             ws.data_ptr(), ws.numel(), int(training), logits.data_ptr(),
                                             _lib.stream_ptr()), "mmu_mmbt_forward")
        saved = (cfg, ws, (tokens, txt, mask, segment, idx, shadow, seed), logits)
        return stamp_workspace(self, ws, saved) if training else saved

    def _engine_backward(self, saved, dlogits, need_dimg):
        cfg, ws, (tokens, txt, mask, segment, idx, shadow, seed), _ = saved
        check_workspace(self, ws, saved)
        self._ensure_grad_views()
        dimg = torch.empty_like(tokens) if need_dimg else None
        inp = _lib.MmbtInputs(txt.data_ptr(), mask.data_ptr(), segment.data_ptr(), tokens.data_ptr(),
                              _lib.ptr(idx), 0 if idx is None else idx.shape[-1],
                              int(idx is not None and idx.dim() == 2), _lib.ptr(shadow), _lib.ptr(dimg), seed)
        _lib.check(_lib.lib.mmu_mmbt_backward(C.byref(cfg), self._flat.data_ptr(), C.byref(inp),
                                              ws.data_ptr(), ws.numel(), dlogits.data_ptr(),
                                              self._flat_grad.data_ptr(), _lib.stream_ptr()),
                   "mmu_mmbt_backward")
        sync = getattr(self, "_flat_sync", None)
        if sync is not None:
            sync.launch(self)  # gradient all-reduce overlaps the image encoder's backward
        return dimg

    # -------------------------------------------------------------- reference protocol
    def _tokens(self, img):
        if img.dim() == 3:
            return img  # pooled image tokens (cached output of a frozen image encoder)
        ie = self.enc._modules["img_encoder"]
        if ie is None:
            raise _lib.MMUError("raw images need an image encoder; pass pooled tokens (B, N, d) instead")
        return ie(img)

    def forward_indices(self, txt, mask, segment, img, indices=None):
        """General entry: ``indices`` (sequence of ints over ``[CLS] img.. [SEP] | text..``, or None
        for all positions) selects what enters the encoder."""
        tokens = self._tokens(img)
        if self.training and torch.is_grad_enabled():
            if tokens.requires_grad:
                return _MmbtForward.apply(tokens, self, txt, mask, segment, indices)
            # frozen / cached image encoder: a parameter is the differentiable input that makes
            # autograd call backward
            anchor = next(p for p in self._param_list if p.requires_grad)
            return _MmbtForwardAnchored.apply(anchor, self, tokens, txt, mask, segment, indices)
        return self._engine_forward(tokens, txt, mask, segment, indices, training=False)[-1]

    @torch.no_grad()
    def forward_index_lists(self, txt, mask, segment, img, index_lists):
        """Several EQUALLY LONG index lists (robustness variants) evaluated in one pass: the batch is
        replicated along the batch axis and every replica gets its own list
        (``mmu_mmbt_inputs.indices_per_sample``).  Inference only.  Returns ``(V, B, C)``; row v is
        what ``forward_indices(..., index_lists[v])`` returns."""
        V, B = len(index_lists), txt.shape[0]
        idx = torch.tensor([list(map(int, il)) for il in index_lists], dtype=torch.int32)
        idx = idx.repeat_interleave(B, 0)
        tokens = self._tokens(img)

        def rep(t):
            return t.repeat(V, *([1] * (t.dim() - 1)))

        logits = self._engine_forward(rep(tokens), rep(txt), rep(mask), rep(segment), idx, training=False)[-1]
        return logits.view(V, B, -1)

    def forward(self, txt, mask, segment, img):
        return self.forward_indices(txt, mask, segment, img, None)

    def forward_img_only(self, txt, mask, segment, img):
        """src/mmbt.py:131-153: only [CLS] img.. [SEP] enter the encoder."""
        return self.forward_indices(txt, mask, segment, img, range(self._n_img + 2))

    def forward_txt_only(self, txt, mask, segment, img):
        """src/mmbt.py:155-184: the image [CLS] row followed by the text."""
        n2 = self._n_img + 2
        return self.forward_indices(txt, mask, segment, img, [0] + list(range(n2, n2 + txt.shape[1])))

    @staticmethod
    def control_indices(total_embeds, num_embeds):
        """The draw of src/mmbt.py:198-201 (host ``torch.randperm``, position 0 always kept)."""
        ind, _ = torch.sort(torch.randperm(total_embeds - 1)[:num_embeds] + 1)
        return [0] + [int(i) for i in ind]

    def forward_control(self, txt, mask, segment, img, control_modal):
        """src/mmbt.py:186-234."""
        total = txt.shape[1] + self._n_img + 2
        if control_modal == "image":
            num = self._n_img + 1
        elif control_modal == "text":
            num = txt.shape[1]
        else:
            raise ValueError("control_modal must be either image or text")
        return self.forward_indices(txt, mask, segment, img, self.control_indices(total, num))

    def compute_loss(self, y_hat, y, eval=False):
        """src/mmbt.py:261-262: ``CrossEntropyLoss()(y_hat, y)`` on (B, C) logits."""
        y_hat = y_hat if y_hat.is_contiguous() else y_hat.contiguous()
        y = y.reshape(-1).contiguous()
        if y_hat.requires_grad:
            return _CELoss.apply(y_hat, y, self)
        N = y_hat.shape[0]
        _, _, _, accum = ops.heads_uncertainty_epilogue(y_hat.detach().view(N, 1, -1), y.view(N, 1), 0)
        self._remember_epilogue(y_hat, 0, accum)
        return _loss_from_accum(accum)


class _MmbtForwardAnchored(torch.autograd.Function):
    """Same as ``_MmbtForward`` when the image tokens carry no gradient (frozen / cached image
    encoder): a parameter serves as the differentiable input that makes autograd call backward."""

    @staticmethod
    def forward(ctx, anchor, model, tokens, txt, mask, segment, indices):
        ctx.model = model
        ctx.saved = model._engine_forward(tokens, txt, mask, segment, indices, training=True)
        return ctx.saved[-1]

    @staticmethod
    def backward(ctx, dlogits):
        ctx.model._engine_backward(ctx.saved, dlogits.contiguous(), False)
        return (None,) * 7
