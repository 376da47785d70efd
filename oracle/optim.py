"""Oracle: optimiser and LR schedule of the FLAVA fusion run (TEST INFRASTRUCTURE).

``train.py:196-210`` builds ``torch.optim.AdamW(lr, betas=(0.9, 0.98), eps=1e-9,
weight_decay=wd)`` and HuggingFace's ``get_cosine_schedule_with_warmup`` stepped once per
batch (``src/framework.py:314-315``).  Both algorithms are published; they are restated here
explicitly so the fused CUDA optimiser can be checked without either library.
"""
import math

import torch


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=1e-3):
    """One decoupled-weight-decay Adam step, PyTorch's formulation:

        p <- p (1 - lr wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
        p <- p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)

    ``step`` is the 1-based step count AFTER incrementing.  Returns new (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = torch.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def cosine_with_warmup_factor(current_step, num_warmup_steps, num_training_steps, num_cycles=0.5):
    """LR multiplier of ``transformers.get_cosine_schedule_with_warmup`` after ``current_step``
    scheduler steps (train.py:204-208: warm-up = 3 epochs of steps)."""
    if current_step < num_warmup_steps:
        return float(current_step) / float(max(1, num_warmup_steps))
    progress = float(current_step - num_warmup_steps) / float(
        max(1, num_training_steps - num_warmup_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


def sgd_momentum_step(p, g, buf, lr, momentum=0.9, weight_decay=1e-3, first=False):
    """``torch.optim.SGD(momentum, weight_decay)`` (train_fashionmnist.py:113-116)."""
    g = g + weight_decay * p
    buf = g.clone() if first else momentum * buf + g
    return p - lr * buf, buf
