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
