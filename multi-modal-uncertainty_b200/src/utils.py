"""Seeding, checkpoint I/O and device helpers (reference ``src/utils.py:14-21, 76-106``)."""
import os
import random

import numpy as np
import torch


def set_seed(seed):
    """Reference src/utils.py:14-21: python, numpy, torch (CPU + CUDA) and cudnn determinism."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def save_weights(model, optimizer, filename):
    """Checkpoint layout of the reference (src/utils.py:98-106): {'model': ..., 'optimizer': ...}."""
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict()}, filename)


def torch_apply(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(torch_apply(o, fn) for o in obj)
    if isinstance(obj, dict):
        return {k: torch_apply(v, fn) for k, v in obj.items()}
    return obj


def torch_to(obj, device, non_blocking=True):
    """Reference src/utils.py:76-90: move every tensor inside a nested container."""
    dev = torch.device("cuda", device) if isinstance(device, int) else device
    return torch_apply(obj, lambda t: t.to(dev, non_blocking=non_blocking))
