"""Fused AdamW over the model's flat parameter buffer + the cosine-with-warm-up schedule.

Drop-in for ``torch.optim.AdamW(model.parameters(), lr, betas=(0.9, 0.98), eps=1e-9,
weight_decay=wd)`` and ``transformers.get_cosine_schedule_with_warmup`` as configured in
reference ``train.py:196-210``.  One CUDA kernel updates every parameter (28 B/param of HBM
traffic); ``param_groups[0]["lr"]`` stays the single source of truth so LR schedulers work.
"""
import math

import torch

from ._backend import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.98), eps=1e-9, weight_decay=1e-3):
        params = list(params)
        owners = {id(getattr(p, "_mmu_owner", None)): getattr(p, "_mmu_owner", None) for p in params}
        if len(owners) != 1 or None in owners.values():
            raise ValueError("FusedAdamW needs the parameters of exactly one mmu_b200 model "
                             "(flat-buffer backed); there is no per-tensor fallback")
        self.model = next(iter(owners.values()))
        if len(params) != len(list(self.model.parameters())):
            raise ValueError("FusedAdamW must own ALL parameters of the model (flat update)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._m = None
        self._v = None
        self.grad_scale = 1.0  # set to 1/world_size by the data-parallel wrapper

    def _moments(self):
        flat = self.model._flat
        if self._m is None or self._m.device != flat.device:
            m = torch.zeros_like(flat)
            v = torch.zeros_like(flat)
            if self._m is not None:
                m.copy_(self._m)
                v.copy_(self._v)
            self._m, self._v = m, v
        return self._m, self._v

    def zero_grad(self, set_to_none: bool = False):
        # gradients are views of one flat buffer: zero it in place, never drop the views
        self.model._flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        ddp = getattr(self.model, "_ddp", None)
        if ddp is not None:
            ddp.wait()
        g = self.param_groups[0]
        m, v = self._moments()
        self._step += 1
        model = self.model
        shadow = model._shadow if model._shadow is not None and \
            model._shadow.device == model._flat.device else None
        ops.adamw_flat_step(model._flat, model._flat_grad, m, v, self._step, float(g["lr"]),
                            betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"],
                            grad_scale=self.grad_scale, p_bf16=shadow)
        if shadow is not None:  # the kernel rewrote the bf16 shadow together with the master
            model._shadow_stamp = model._param_stamp()

    # checkpoint format compatible with torch.optim.AdamW's (per-parameter state, reference
    # src/utils.py:98-106 saves {'model', 'optimizer'})
    def state_dict(self):
        m, v = self._moments()
        state = {}
        for i, p in enumerate(self.param_groups[0]["params"]):
            off = p.data_ptr() - self.model._flat.data_ptr()
            off //= 4
            n = p.numel()
            state[i] = {"step": torch.tensor(float(self._step)),
                        "exp_avg": m[off:off + n].view(p.shape).clone(),
                        "exp_avg_sq": v[off:off + n].view(p.shape).clone()}
        group = {k: val for k, val in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.param_groups[0]["params"])))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        m, v = self._moments()
        for k, val in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = val
        for i, p in enumerate(self.param_groups[0]["params"]):
            st = sd["state"].get(i)
            if st is None:
                continue
            off = (p.data_ptr() - self.model._flat.data_ptr()) // 4
            n = p.numel()
            m[off:off + n].view(p.shape).copy_(st["exp_avg"])
            v[off:off + n].view(p.shape).copy_(st["exp_avg_sq"])
            self._step = int(float(st["step"]))


def warmup_linear(x, warmup=0.002):
    """pytorch_pretrained_bert.optimization.warmup_linear (the ``schedule`` BertAdam defaults to)."""
    if x < warmup:
        return x / warmup
    return 1.0 - x


class BertAdam(torch.optim.Optimizer):
    """Drop-in for ``pytorch_pretrained_bert.BertAdam`` as the reference configures it for MMBT
    (``train.py:136-147``: two parameter groups, weight decay 0.01 / 0, ``lr``, ``warmup``,
    ``t_total``), as ONE fused update per flat buffer (``mmu_bertadam_flat_step``): per-tensor
    ``clip_grad_norm_(p, max_grad_norm)``, Adam moments without bias correction, weight decay added
    to the update, ``lr * warmup_linear(step / t_total, warmup)`` with ``state['step']`` kept per
    tensor.  Parameters with ``requires_grad == False`` are skipped, as tensors whose ``.grad`` is
    ``None`` are in the reference (the frozen epochs of ``src/framework.py:246-285``).  Every
    parameter must be flat-buffer backed (an mmu_b200 model); there is no per-tensor fallback."""

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999,
                 e=1e-6, weight_decay=0.01, max_grad_norm=1.0):
        if schedule != "warmup_linear":
            raise ValueError("only the 'warmup_linear' schedule (the reference's) is implemented")
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._owners = []
        for g in self.param_groups:
            for p in g["params"]:
                owner = getattr(p, "_mmu_owner", None)
                if owner is None:
                    raise ValueError("BertAdam needs flat-buffer backed parameters of mmu_b200 models")
                if not any(owner is o for o in self._owners):
                    self._owners.append(owner)
        self._mv = {}
        self._seg_cache = {}
        self._steps = {}
        self.grad_scale = 1.0

    def _moments(self, owner):
        flat = owner._flat
        mv = self._mv.get(id(owner))
        if mv is None or mv[0].device != flat.device:
            new = (torch.zeros_like(flat), torch.zeros_like(flat))
            if mv is not None:
                new[0].copy_(mv[0])
                new[1].copy_(mv[1])
            mv = self._mv[id(owner)] = new
        return mv

    def zero_grad(self, set_to_none: bool = False):
        for o in self._owners:
            o._flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        import ctypes as C  # noqa: F401
        from ._backend import _lib
        sync = getattr(self, "_flat_sync", None)
        if sync is not None:
            sync.all_reduce_grads()  # sum over ranks; 1/world is folded into the kernel (grad_scale)
        for owner in self._owners:
            flat = owner._flat
            active = [(p, g) for g in self.param_groups for p in g["params"]
                      if getattr(p, "_mmu_owner", None) is owner and p.requires_grad]
            if not active:
                continue
            key = (id(owner), flat.data_ptr(), tuple(id(p) for p, _ in active))
            cached = self._seg_cache.get(key)
            if cached is None:
                segs = torch.tensor([[(p.data_ptr() - flat.data_ptr()) // 4, p.numel()] for p, _ in active],
                                    dtype=torch.int64, device=flat.device)
                # pinned staging of the per-tensor (weight decay, lr) rows: a RING of buffers, each
                # guarded by an event recorded behind its upload, so that a host that runs ahead of
                # the GPU (gradient accumulation, no .item() per step) never rewrites rows whose DMA
                # is still pending
                host = _PinnedRing((len(active), 2))
                cached = self._seg_cache[key] = (
                    segs, host, torch.empty(len(active), 2, dtype=torch.float32, device=flat.device),
                    torch.empty(len(active), dtype=torch.float32, device=flat.device),
                    max(p.numel() for p, _ in active))
            segs, ring, hyper, norms, max_numel = cached
            host = ring.acquire()
            for i, (p, g) in enumerate(active):
                st = self._steps.get(id(p), 0)
                if g["t_total"] != -1:
                    lr = g["lr"] * warmup_linear(st / g["t_total"], g["warmup"])
                else:
                    lr = g["lr"]
                host[i, 0] = g["weight_decay"]
                host[i, 1] = lr
                self._steps[id(p)] = st + 1
            hyper.copy_(host, non_blocking=True)
            ring.release()
            g0 = self.param_groups[0]
            m, v = self._moments(owner)
            shadow = getattr(owner, "_shadow", None)
            if shadow is not None and shadow.device != flat.device:
                shadow = None
            _lib.check(_lib.lib.mmu_bertadam_flat_step(
                flat.data_ptr(), owner._flat_grad.data_ptr(), m.data_ptr(), v.data_ptr(), _lib.ptr(shadow),
                segs.data_ptr(), hyper.data_ptr(), norms.data_ptr(), len(active), max_numel,
                g0["b1"], g0["b2"], g0["e"], g0["max_grad_norm"], float(self.grad_scale),
                _lib.stream_ptr()),
                "mmu_bertadam_flat_step")
            if shadow is not None:
                if len(active) == len(list(owner._param_list)):
                    owner._shadow_stamp = owner._param_stamp()
                else:
                    owner.invalidate_shadow()

    def state_dict(self):
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                owner = p._mmu_owner
                m, v = self._moments(owner)
                off = (p.data_ptr() - owner._flat.data_ptr()) // 4
                n = p.numel()
                state[idx] = {"step": self._steps.get(id(p), 0),
                              "next_m": m[off:off + n].view(p.shape).clone(),
                              "next_v": v[off:off + n].view(p.shape).clone()}
                ids.append(idx)
                idx += 1
            gg = {k: val for k, val in g.items() if k != "params"}
            gg["params"] = ids
            groups.append(gg)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        idx = 0
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, val in sg.items():
                if k != "params":
                    g[k] = val
            for p in g["params"]:
                st = sd["state"].get(idx)
                idx += 1
                if st is None:
                    continue
                owner = p._mmu_owner
                m, v = self._moments(owner)
                off = (p.data_ptr() - owner._flat.data_ptr()) // 4
                n = p.numel()
                m[off:off + n].view(p.shape).copy_(st["next_m"])
                v[off:off + n].view(p.shape).copy_(st["next_v"])
                self._steps[id(p)] = int(st["step"])


class _PinnedRing:
    """``slots`` pinned host buffers handed out round-robin; ``release()`` records a CUDA event on
    the current stream behind the asynchronous upload that read the buffer, ``acquire()`` waits for
    that event before the buffer is rewritten."""

    def __init__(self, shape, dtype=torch.float32, slots=4):
        self.bufs = [torch.empty(shape, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.k = -1

    def acquire(self):
        self.k = (self.k + 1) % len(self.bufs)
        ev = self.events[self.k]
        if ev is not None:
            ev.synchronize()
        return self.bufs[self.k]

    def release(self):
        ev = self.events[self.k]
        if ev is None:
            ev = self.events[self.k] = torch.cuda.Event()
        ev.record()


def cosine_with_warmup_lambda(num_warmup_steps, num_training_steps, num_cycles=0.5):
    def fn(current_step):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        progress = float(current_step - num_warmup_steps) / float(
            max(1, num_training_steps - num_warmup_steps))
        return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))
    return fn


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps,
                                    num_cycles=0.5, last_epoch=-1):
    """Same multiplier as ``transformers.get_cosine_schedule_with_warmup`` (train.py:204-208),
    stepped once per batch by ``Model_`` (``scheduler_step_on='batch'``)."""
    return torch.optim.lr_scheduler.LambdaLR(
        optimizer, cosine_with_warmup_lambda(num_warmup_steps, num_training_steps, num_cycles),
        last_epoch)
