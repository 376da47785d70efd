"""Deterministic, library-independent parameters for fixtures whose tensors are too large to
store: the golden script and the tests regenerate the SAME values from a seed (CPU
``torch.Generator``), so only outputs need to live in the fixture."""
import math

import torch


def det_image_encoder_state(shapes, seed):
    """``shapes``: {state_dict key: shape} of an ImageEncoder (reference src/mmbt.py:15-45)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif len(shape) == 4:
            sd[name] = torch.randn(*shape, generator=g) * math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
        elif name.endswith("running_var"):
            sd[name] = 1.0 + 0.2 * torch.rand(*shape, generator=g)
        elif name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(*shape, generator=g)
        elif name.endswith(".weight"):
            sd[name] = 1.0 + 0.2 * torch.randn(*shape, generator=g)
        else:
            sd[name] = 0.2 * torch.randn(*shape, generator=g)
    return sd


def grad_digest(g, limit=65536, n=4096):
    """Full tensor when small, else an evenly strided sample of n elements + the L2 norm."""
    if g.numel() <= limit:
        return {"full": g.detach().clone()}
    stride = g.numel() // n
    return {"stride": stride, "sample": g.detach().flatten()[::stride][:n].clone(),
            "norm": float(g.detach().double().norm())}


def digest_error(digest, g):
    """max |g - golden| / max |golden| over what the digest holds (and the norm's relative error)."""
    g = g.detach().double().cpu()
    if "full" in digest:
        ref = digest["full"].double()
        return float((g - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    ref = digest["sample"].double()
    got = g.flatten()[::digest["stride"]][:ref.numel()]
    e = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    return max(e, abs(float(g.norm()) - digest["norm"]) / max(digest["norm"], 1e-30))
