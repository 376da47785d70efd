"""Oracle: the four-view FashionMNIST ResNet (TEST INFRASTRUCTURE, never imported by the product).

Restates reference ``src/model.py:17-100`` (``ResNet`` / ``MultiHeadFC`` / ``MIMOResNet``) and
``src/layers.py:7-38`` (``BasicBlock``) in explicit tensor arithmetic: a convolution is an
unfold (im2col) followed by one matrix product, BatchNorm is written out with its batch /
running statistics, pooling is a slice mean.  Parameters and buffers are a plain ``dict`` keyed
by the reference's ``state_dict`` names.  Works in the dtype of its inputs.
"""
import torch

from .fusion import compute_loss, linear


def conv2d(x, weight, stride=1, padding=1):
    """``nn.Conv2d(bias=False)``: x (B, Ci, H, W), weight (Co, Ci, k, k) -> (B, Co, Ho, Wo)."""
    B, Ci, H, W = x.shape
    Co, _, k, _ = weight.shape
    Ho = (H + 2 * padding - k) // stride + 1
    Wo = (W + 2 * padding - k) // stride + 1
    cols = torch.nn.functional.unfold(x, kernel_size=k, padding=padding, stride=stride)  # (B, Ci*k*k, Ho*Wo)
    out = weight.reshape(Co, -1) @ cols  # (B, Co, Ho*Wo)
    return out.reshape(B, Co, Ho, Wo)


def batch_norm(x, P, prefix, training, buffers_out=None, momentum=0.1, eps=1e-5):
    """``nn.BatchNorm2d``: batch statistics (biased variance) in training, running statistics in
    eval; in training the running buffers move by ``momentum`` towards the batch mean and the
    UNBIASED batch variance (torch semantics) -- returned through ``buffers_out``."""
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        if buffers_out is not None:
            n = x.numel() // x.shape[1]
            buffers_out[prefix + ".running_mean"] = (1 - momentum) * P[prefix + ".running_mean"] + momentum * mean.detach()
            buffers_out[prefix + ".running_var"] = (1 - momentum) * P[prefix + ".running_var"] + \
                momentum * var.detach() * n / max(n - 1, 1)
            buffers_out[prefix + ".num_batches_tracked"] = P[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    xh = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
    return xh * w[None, :, None, None] + b[None, :, None, None]


def basic_block(x, P, prefix, stride, training, buffers_out):
    """src/layers.py:21-38."""
    out = conv2d(x, P[prefix + ".conv1.weight"], stride, 1)
    out = torch.relu(batch_norm(out, P, prefix + ".bn1", training, buffers_out))
    out = conv2d(out, P[prefix + ".conv2.weight"], 1, 1)
    out = batch_norm(out, P, prefix + ".bn2", training, buffers_out)
    residual = x
    if prefix + ".downsample.0.weight" in P:
        residual = conv2d(x, P[prefix + ".downsample.0.weight"], stride, 0)
        residual = batch_norm(residual, P, prefix + ".downsample.1", training, buffers_out)
    return torch.relu(out + residual)


def mimo_resnet_forward(P, x, num_classes, training=True, buffers_out=None):
    """``MIMOResNet.forward`` (src/model.py:81-100): views -> channels, conv/bn/relu stem, two
    stages of two BasicBlocks (the second strided), ``AvgPool2d(4)`` (one 4x4 window of the 7x7
    map), ``MultiHeadFC`` (src/model.py:58-70).  Returns (B, E, C)."""
    if x.dim() == 5:
        x = x.reshape(x.shape[0], -1, x.shape[3], x.shape[4])
    h = conv2d(x, P["conv1.weight"], 1, 1)
    h = torch.relu(batch_norm(h, P, "bn1", training, buffers_out))
    h = basic_block(h, P, "layer1.0", 1, training, buffers_out)
    h = basic_block(h, P, "layer1.1", 1, training, buffers_out)
    h = basic_block(h, P, "layer2.0", 2, training, buffers_out)
    h = basic_block(h, P, "layer2.1", 1, training, buffers_out)
    pooled = h[:, :, :4, :4].mean(dim=(2, 3))  # AvgPool2d(4) on 7x7: the top-left window only
    out = linear(pooled, P["output_layer.fc.weight"], P["output_layer.fc.bias"])
    return out.reshape(out.shape[0], -1, num_classes)


def loss_and_grads(P, x, y, num_classes):
    """Train-mode forward + CE + backward via autograd over the explicit forward (checker only).
    Returns (logits, loss, grads, updated buffers)."""
    is_param = lambda k: not (k.endswith("running_mean") or k.endswith("running_var") or
                              k.endswith("num_batches_tracked"))
    leaves = {k: (v.detach().clone().requires_grad_(True) if is_param(k) else v) for k, v in P.items()}
    buffers = {}
    logits = mimo_resnet_forward(leaves, x, num_classes, True, buffers)
    loss = compute_loss(logits, y, eval=False)
    names = [k for k in leaves if is_param(k)]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return logits.detach(), loss.detach(), dict(zip(names, grads)), buffers
