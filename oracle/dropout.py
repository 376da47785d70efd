"""Oracle: the counter-based dropout masks of the CUDA path, restated in integer numpy arithmetic
(TEST INFRASTRUCTURE).

The reference draws its dropout masks from torch's Philox stream (``nn.Dropout`` between ``c_fc``
and QuickGELU, src/model.py:195-201; ``ImageBertEmbeddings.dropout``, src/mmbt.py:56,82; the hidden
and attention-probability dropouts inside pytorch_pretrained_bert's ``BertModel``).  A different
generator cannot reproduce those draws, so parity for p > 0 is twofold: (1) STATISTICAL against
``nn.Dropout``'s definition -- each element kept independently with probability 1 - p, survivors
scaled by 1 / (1 - p); (2) BIT-EXACT against this restatement of the engine's own mask function
(``csrc/dropout.cuh``), which lets the oracle apply the very same masks and hold the arithmetic
around them to the usual 1e-3.

    site_seed = splitmix64(seed + 0x9E3779B97F4A7C15 * (site + 1))
    h = idx * 0x9E3779B1 + lo;  h ^= h >> 16;  h *= 0x7FEB352D;  h ^= h >> 15
    h += hi;                    h *= 0x846CA68B;  h ^= h >> 16          (mod 2^32)
    keep  <=>  h >= floor(p * 2^32)
"""
import numpy as np
import torch

M64 = (1 << 64) - 1
M32 = np.uint64(0xFFFFFFFF)


def site_seed(seed, site):
    z = (int(seed) + 0x9E3779B97F4A7C15 * (int(site) + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def threshold(p):
    p = float(np.float32(p))
    if not p > 0.0:
        return 0
    return min(int(p * 4294967296.0), 4294967295)


def hash32(idx, lo, hi):
    h = (idx.astype(np.uint64) * np.uint64(0x9E3779B1) + np.uint64(lo)) & M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x7FEB352D)) & M32
    h ^= h >> np.uint64(15)
    h = (h + np.uint64(hi)) & M32
    h = (h * np.uint64(0x846CA68B)) & M32
    h ^= h >> np.uint64(16)
    return h


def keep_mask(p, seed, site, numel, shape=None, offset=0):
    """Boolean keep mask of the ``numel`` elements (row-major counter idx = offset + 0..numel-1)
    that dropout site ``site`` sees under ``seed``; all True for p == 0."""
    idx = (np.arange(numel, dtype=np.uint64) + np.uint64(offset)) & M32
    ss = site_seed(seed, site)
    h = hash32(idx, ss & 0xFFFFFFFF, ss >> 32)
    keep = torch.from_numpy(h >= np.uint64(threshold(p)))
    return keep.reshape(shape) if shape is not None else keep


def multiplier(p, seed, site, numel, shape=None, dtype=torch.float32, offset=0):
    """0 or 1 / (1 - p) per element: what ``nn.Dropout(p)`` multiplies by in training mode.  The
    scale is the fp32 value 1.0f / (1.0f - p) the kernels use."""
    keep = keep_mask(p, seed, site, numel, shape, offset)
    scale = float(np.float32(1.0) / (np.float32(1.0) - np.float32(p))) if p > 0 else 1.0
    return keep.to(dtype) * scale
