"""Oracle: forward pass, losses and metrics of the fusion models (TEST INFRASTRUCTURE).

Restates reference ``src/model.py`` in explicit tensor arithmetic.  Parameters are passed as a
plain ``dict`` keyed by the reference's ``state_dict`` names, so a reference checkpoint can be
fed in unchanged.  Every function works in the dtype of its inputs (fp32 to mirror the
reference, fp64 to serve as ground truth).
"""
import math

import torch


# --------------------------------------------------------------------------- primitives
def linear(x, weight, bias=None):
    """``nn.Linear``: y = x W^T + b with W stored (out, in)."""
    y = x @ weight.transpose(0, 1)
    return y if bias is None else y + bias


def layer_norm(x, weight, bias, eps=1e-5):
    """``nn.LayerNorm`` over the last axis, biased variance (src/model.py:174-180, :252-253)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc / torch.sqrt(var + eps) * weight + bias


def quick_gelu(x):
    """src/model.py:183-185."""
    return x * (1.0 / (1.0 + torch.exp(-1.702 * x)))


def softmax_lastdim(s):
    m = s.max(dim=-1, keepdim=True).values
    e = torch.exp(s - m)
    return e / e.sum(dim=-1, keepdim=True)


def batch_axis_attention(x, in_w, in_b, out_w, out_b, n_head):
    """``nn.MultiheadAttention`` fed (B, L, D) with batch_first=False (src/model.py:193,205-207).

    PyTorch therefore treats axis 0 (the mini-batch) as the sequence and axis 1 (tokens) as the
    batch: token position l of sample b attends to token position l of every other sample.
    """
    B, L, D = x.shape
    hd = D // n_head
    qkv = linear(x, in_w, in_b)  # (B, L, 3D), packed q | k | v
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]

    def split_heads(t):  # (B, L, D) -> (L, H, B, hd)
        return t.reshape(B, L, n_head, hd).permute(1, 2, 0, 3)

    q, k, v = split_heads(q), split_heads(k), split_heads(v)
    s = (q * (1.0 / math.sqrt(hd))) @ k.transpose(-1, -2)  # (L, H, B, B)
    p = softmax_lastdim(s)
    o = p @ v  # (L, H, B, hd)
    o = o.permute(2, 0, 1, 3).reshape(B, L, D)
    return linear(o, out_w, out_b)


def residual_attention_block(x, P, pre, n_head, drop_mult=None):
    """src/model.py:188-212.  Effective MLP order is c_fc -> dropout -> QuickGELU -> c_proj
    (duplicate OrderedDict key).  ``drop_mult`` (training with p > 0):
    the (B, L, 4D) tensor of 0 / (1 / (1 - p)) factors nn.Dropout multiplies the c_fc output by."""
    h = layer_norm(x, P[pre + "ln_1.weight"], P[pre + "ln_1.bias"])
    x = x + batch_axis_attention(h, P[pre + "attn.in_proj_weight"], P[pre + "attn.in_proj_bias"],
                                 P[pre + "attn.out_proj.weight"], P[pre + "attn.out_proj.bias"],
                                 n_head)
    h = layer_norm(x, P[pre + "ln_2.weight"], P[pre + "ln_2.bias"])
    z = linear(h, P[pre + "mlp.c_fc.weight"], P[pre + "mlp.c_fc.bias"])
    if drop_mult is not None:
        z = z * drop_mult
    u = quick_gelu(z)
    return x + linear(u, P[pre + "mlp.c_proj.weight"], P[pre + "mlp.c_proj.bias"])


def transformer(x, P, pre, n_layers, n_head, dropout=None):
    """``dropout = (p, seed)``: training-mode masks of the engine's counter-based generator
    (oracle/dropout.py; site = layer index, element counter = row-major index of the engine's
    POSITION-major [L*B, 4D] activation: (l * B + b) * 4D + column)."""
    for i in range(n_layers):
        mult = None
        if dropout is not None and dropout[0] > 0:
            from . import dropout as _d
            B, L, D = x.shape
            mult = _d.multiplier(dropout[0], dropout[1], i, B * L * 4 * D, (L, B, 4 * D), x.dtype).permute(1, 0, 2)
        x = residual_attention_block(x, P, f"{pre}resblocks.{i}.", n_head, mult)
    return x


def count_layers(P, pre="mm_encoder."):
    n = 0
    while f"{pre}resblocks.{n}.ln_1.weight" in P:
        n += 1
    return n


def count_heads(P):
    n = 0
    while f"output_layers.{n}.weight" in P:
        n += 1
    return n


# ------------------------------------------------------------------------------ models
def flava_fusion_forward(P, x, n_head, avg_pool=False, dropout=None):
    """``FlavaFusionTransfomer.forward`` (src/model.py:258-291) and, when ``class_embeddings``
    is present in ``P``, ``FlavaFusionTransfomerwithCLSToken.forward`` (:330-361).

    A ``None`` modality is dropped from the sequence -- the behaviour the CLS variant
    implements (:334-344) and the robustness script relies on
    (eval_transformer_robustness.py:107-121); the non-CLS reference class dereferences
    ``.shape`` before its ``None`` checks (:266) and cannot run that case as committed.
    """
    img, txt = x
    parts = []
    l_img = l_txt = 0
    if img is not None:
        fi = linear(img, P["image_to_mm_projection.weight"], P["image_to_mm_projection.bias"])
        l_img = fi.shape[1]
        parts.append(fi)
    if txt is not None:
        ft = linear(txt, P["text_to_mm_projection.weight"], P["text_to_mm_projection.bias"])
        l_txt = ft.shape[1]
        parts.append(ft)
    mm = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    if "class_embeddings" in P:
        cls = P["class_embeddings"].transpose(0, 1).unsqueeze(0).expand(mm.shape[0], -1, -1)
        mm = torch.cat([cls, mm], dim=1)
        avg_pool = False  # the CLS variant never pools (src/model.py:354-357)
    mm = layer_norm(mm, P["ln_pre.weight"], P["ln_pre.bias"])
    out = transformer(mm, P, "mm_encoder.", count_layers(P), n_head, dropout)
    out = layer_norm(out, P["ln_post.weight"], P["ln_post.bias"])
    heads = []
    if avg_pool:
        heads.append(linear(out[:, :l_img, :].mean(1), P["output_layers.0.weight"],
                            P["output_layers.0.bias"]))
        heads.append(linear(out[:, l_img:l_img + l_txt, :].mean(1), P["output_layers.1.weight"],
                            P["output_layers.1.bias"]))
    else:
        for i in range(count_heads(P)):
            heads.append(linear(out[:, i, :], P[f"output_layers.{i}.weight"],
                                P[f"output_layers.{i}.bias"]))
    return torch.stack(heads, dim=1)  # (B, E, C)


def mimo_transformer_forward(P, x, n_head):
    """``MIMOTransfomer.forward`` (src/model.py:138-159)."""
    b, e, c, h, w = x.shape
    t = x.reshape(b, e * c, h * w)
    t = linear(t, P["image_to_mm_projection.weight"], P["image_to_mm_projection.bias"])
    t = layer_norm(t, P["ln_pre.weight"], P["ln_pre.bias"])
    t = transformer(t, P, "mm_encoder.", count_layers(P), n_head)
    t = layer_norm(t, P["ln_post.weight"], P["ln_post.bias"])
    t = t.reshape(b, e, c, -1).mean(2)
    heads = [linear(t[:, i, :], P[f"output_layers.{i}.weight"], P[f"output_layers.{i}.bias"])
             for i in range(count_heads(P))]
    return torch.stack(heads, dim=1)


def multi_head_fc(x, weight, bias, num_classes):
    """``MultiHeadFC.forward`` (src/model.py:58-70): one Linear, split into E heads."""
    out = linear(x, weight, bias)
    B = out.shape[0]
    return out.reshape(B, -1, num_classes)


# --------------------------------------------------------------------- loss and metric
def cross_entropy_mean(logits, y):
    """``nn.CrossEntropyLoss()`` (mean reduction): mean_i [logsumexp(z_i) - z_i[y_i]]."""
    m = logits.max(dim=1, keepdim=True).values
    lse = m.squeeze(1) + torch.log(torch.exp(logits - m).sum(dim=1))
    picked = logits.gather(1, y.view(-1, 1)).squeeze(1)
    return (lse - picked).mean()


def compute_loss(y_hat, y, eval=False):
    """src/model.py:293-304 (identical in :102-112, :161-171, :363-374)."""
    assert y.shape[0] == y_hat.shape[0]
    y = y.reshape(-1)
    if not eval:
        y_hat = y_hat.reshape(-1, y_hat.shape[2])  # one CE term per ensemble member
    else:
        y_hat = y_hat.mean(1)  # CE on the head-mean of the LOGITS
    return cross_entropy_mean(y_hat, y)


def acc(y_pred, y_true, eval, dummy_dim=False):
    """train.py:119-130.  argmax takes the first maximal index, as ``Tensor.max(1)`` on CPU."""
    if dummy_dim:
        if not eval:
            y_pred = y_pred.reshape(-1, y_pred.shape[2])
            y_true = y_true.reshape(-1)
        else:
            y_pred = y_pred.mean(1)
    pred = y_pred.argmax(dim=1)
    return (pred == y_true).to(torch.float32).mean() * 100


def predictions(y_pred, eval):
    """Integer predictions behind ``acc`` (for bit-exact comparison)."""
    if not eval:
        y_pred = y_pred.reshape(-1, y_pred.shape[2])
    else:
        y_pred = y_pred.mean(1)
    return y_pred.argmax(dim=1)


# -------------------------------------------------------------------------- train step
def loss_and_grads(P, x, y, n_head, avg_pool=False, model="flava", dropout=None):
    """Forward + CE + backward.  Gradients come from autograd over the explicit forward above
    (this is the checker, not the product).  Returns (logits, loss, {name: grad})."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    if model == "flava":
        logits = flava_fusion_forward(leaves, x, n_head, avg_pool, dropout)
    else:
        logits = mimo_transformer_forward(leaves, x, n_head)
    loss = compute_loss(logits, y, eval=False)
    names = list(leaves)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
    return logits.detach(), loss.detach(), gd
