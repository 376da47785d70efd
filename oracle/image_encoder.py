"""Oracle: the MMBT image encoder (TEST INFRASTRUCTURE, never imported by the product).

Restates reference ``src/mmbt.py:15-45`` (``ImageEncoder``: torchvision ResNet-152 trunk
``children()[:-2]`` + ``AdaptiveAvgPool2d`` / ``AdaptiveMaxPool2d`` + flatten/transpose) in explicit
tensor arithmetic over a ``dict`` keyed by the reference's ``state_dict`` names (``model.0.weight``,
``model.4.0.conv1.weight`` ...).  The trunk is torchvision's ``Bottleneck`` ResNet v1.5 (stride on
the 3x3 convolution): convolution = unfold + matmul and BatchNorm as in ``oracle/resnet.py``.
Pinned by ``tests/test_oracle_golden.py`` to goldens from the unmodified reference class.
"""
import torch

from .resnet import batch_norm, conv2d


def max_pool_3x3_s2_p1(x):
    """``nn.MaxPool2d(kernel_size=3, stride=2, padding=1)``."""
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    xp = torch.full((B, C, H + 2, W + 2), float("-inf"), dtype=x.dtype)
    xp[:, :, 1:H + 1, 1:W + 1] = x
    taps = [xp[:, :, ky:ky + 2 * Ho:2, kx:kx + 2 * Wo:2] for ky in range(3) for kx in range(3)]
    return torch.stack(taps, 0).max(0)[0]


def adaptive_pool(x, ph, pw, is_max):
    """``nn.AdaptiveAvgPool2d((ph, pw))`` / ``nn.AdaptiveMaxPool2d``: cell (i, j) covers rows
    [floor(i H / ph), ceil((i + 1) H / ph)) and the analogous columns."""
    B, C, H, W = x.shape
    rows = []
    for i in range(ph):
        y0, y1 = (i * H) // ph, -((-(i + 1) * H) // ph)
        cells = []
        for j in range(pw):
            x0, x1 = (j * W) // pw, -((-(j + 1) * W) // pw)
            win = x[:, :, y0:y1, x0:x1].reshape(B, C, -1)
            cells.append(win.max(-1)[0] if is_max else win.mean(-1))
        rows.append(torch.stack(cells, -1))
    return torch.stack(rows, -2)  # (B, C, ph, pw)


def bottleneck(x, P, prefix, stride, training, buffers_out):
    out = torch.relu(batch_norm(conv2d(x, P[prefix + ".conv1.weight"], 1, 0), P, prefix + ".bn1", training, buffers_out))
    out = torch.relu(batch_norm(conv2d(out, P[prefix + ".conv2.weight"], stride, 1), P, prefix + ".bn2", training, buffers_out))
    out = batch_norm(conv2d(out, P[prefix + ".conv3.weight"], 1, 0), P, prefix + ".bn3", training, buffers_out)
    identity = x
    if prefix + ".downsample.0.weight" in P:
        identity = batch_norm(conv2d(x, P[prefix + ".downsample.0.weight"], stride, 0), P,
                              prefix + ".downsample.1", training, buffers_out)
    return torch.relu(out + identity)


def image_encoder_forward(P, x, layers, pool, is_max, training=False, buffers_out=None):
    """``ImageEncoder.forward``: (B, 3, H, W) -> (B, ph*pw, 2048)."""
    h = torch.relu(batch_norm(conv2d(x, P["model.0.weight"], 2, 3), P, "model.1", training, buffers_out))
    h = max_pool_3x3_s2_p1(h)
    for li, n in enumerate(layers):
        for bi in range(n):
            h = bottleneck(h, P, f"model.{4 + li}.{bi}", 2 if (bi == 0 and li > 0) else 1, training, buffers_out)
    out = adaptive_pool(h, pool[0], pool[1], is_max)
    return out.flatten(2).transpose(1, 2).contiguous()


def tokens_and_grads(P, x, dtokens, layers, pool, is_max):
    """Train-mode forward and the parameter gradients of sum(tokens * dtokens)."""
    is_param = lambda k: not (k.endswith("running_mean") or k.endswith("running_var") or
                              k.endswith("num_batches_tracked"))
    leaves = {k: (v.detach().clone().requires_grad_(True) if is_param(k) else v) for k, v in P.items()}
    buffers = {}
    tok = image_encoder_forward(leaves, x, layers, pool, is_max, True, buffers)
    names = [k for k in leaves if is_param(k)]
    grads = torch.autograd.grad((tok * dtokens).sum(), [leaves[k] for k in names])
    return tok.detach(), dict(zip(names, grads)), buffers
