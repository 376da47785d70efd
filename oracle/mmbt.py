"""Oracle: the MMBT path (TEST INFRASTRUCTURE, never imported by the product).

Restates reference ``src/mmbt.py`` in explicit tensor arithmetic over a plain ``dict`` of
parameters keyed by the reference's ``state_dict`` names:

* ``image_bert_embeddings``  -- ``ImageBertEmbeddings.forward`` (``src/mmbt.py:57-84``)
* ``encoder_input_and_mask`` -- the sequence / mask assembly of ``MultimodalBertEncoder.forward``
  (``:98-122``), ``forward_img_only`` (``:131-148``), ``forward_txt_only`` (``:155-179``) and
  ``forward_control`` (``:186-228``), expressed as ONE index list over the full sequence
* ``control_indices``        -- the index sampling of ``forward_control`` (``:198-201``)
* ``forward`` / ``loss_and_grads`` -- ``MultimodalBertClf.forward*`` + ``compute_loss`` (``:238-262``)

and, for the arithmetic that lives in the absent third-party ``pytorch_pretrained_bert``
(see ``oracle/bert_restated.py``): ``bert_text_embeddings``, ``bert_layer``, ``bert_pooler``,
``bertadam_step``.  Works in the dtype of its inputs (fp64 for pinning, fp32 for timing).
Pinned by ``tests/test_oracle_golden.py`` to goldens produced by the UNMODIFIED ``src/mmbt.py``
running on ``bert_restated`` (``tests/golden/make_golden.py::mmbt_case``).
"""
import math

import torch

from .fusion import linear


def layer_norm(x, w, b, eps=1e-12):
    u = x.mean(-1, keepdim=True)
    s = ((x - u) ** 2).mean(-1, keepdim=True)
    return w * ((x - u) / torch.sqrt(s + eps)) + b


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def bert_text_embeddings(P, ids, segment):
    """BertEmbeddings.forward before its LayerNorm (shared with the image side)."""
    pre = "enc.txt_embeddings."
    pos = torch.arange(ids.shape[1])
    return (P[pre + "word_embeddings.weight"][ids] + P[pre + "position_embeddings.weight"][pos][None]
            + P[pre + "token_type_embeddings.weight"][segment])


def image_bert_embeddings(P, img_tokens, cls_id, sep_id):
    """src/mmbt.py:57-81 before the LayerNorm: [CLS] | Linear(img tokens) | [SEP], positions
    0..N+1, token type 0."""
    pre = "enc.txt_embeddings."
    B, N, _ = img_tokens.shape
    word = P[pre + "word_embeddings.weight"]
    proj = linear(img_tokens, P["enc.img_embeddings.img_embeddings.weight"],
                  P["enc.img_embeddings.img_embeddings.bias"])
    tok = torch.cat([word[cls_id].expand(B, 1, -1), proj, word[sep_id].expand(B, 1, -1)], dim=1)
    pos = P[pre + "position_embeddings.weight"][torch.arange(N + 2)][None]
    typ = P[pre + "token_type_embeddings.weight"][0][None, None]
    return tok + pos + typ


def control_indices(total_embeds, num_embeds, generator=None):
    """src/mmbt.py:198-201: keep position 0, plus a sorted sample of ``num_embeds`` of the others."""
    ind = torch.sort(torch.randperm(total_embeds - 1, generator=generator)[:num_embeds] + 1)[0]
    return torch.cat([torch.zeros(1, dtype=torch.long), ind.long()])


def mode_indices(mode, n_img, s_txt):
    """Index list over [CLS img.. SEP | text..] equivalent to each reference entry point."""
    full = n_img + 2 + s_txt
    if mode == "full":
        return torch.arange(full)
    if mode == "img_only":
        return torch.arange(n_img + 2)
    if mode == "txt_only":
        return torch.cat([torch.zeros(1, dtype=torch.long), torch.arange(n_img + 2, full)])
    raise ValueError(mode)


def _mult(dropout, key, site, like):
    """0 / (1 / (1 - p)) factors of dropout site ``site`` over the row-major elements of ``like``
    (the engine's counter-based masks, oracle/dropout.py); None when that dropout is off."""
    if dropout is None or not dropout.get(key, 0.0) > 0.0:
        return None
    from . import dropout as _d
    return _d.multiplier(dropout[key], dropout["seed"], site, like.numel(), tuple(like.shape), like.dtype)


def bert_layer(P, i, h, ext_mask, n_head, dropout=None):
    """BertLayer (pytorch_pretrained_bert): self-attention with attention-probability dropout,
    BertSelfOutput / BertOutput with hidden dropout before the residual add (train mode only;
    ``dropout = dict(hidden=, attn=, img=, seed=)``, sites 4i+1 / 4i+2 / 4i+3)."""
    pre = f"enc.encoder.layer.{i}."
    B, S, D = h.shape
    hd = D // n_head

    def heads(t):
        return t.reshape(B, S, n_head, hd).permute(0, 2, 1, 3)

    q = heads(linear(h, P[pre + "attention.self.query.weight"], P[pre + "attention.self.query.bias"]))
    k = heads(linear(h, P[pre + "attention.self.key.weight"], P[pre + "attention.self.key.bias"]))
    v = heads(linear(h, P[pre + "attention.self.value.weight"], P[pre + "attention.self.value.bias"]))
    scores = q @ k.transpose(-1, -2) / math.sqrt(hd) + ext_mask
    scores = scores - scores.max(dim=-1, keepdim=True)[0]
    e = torch.exp(scores)
    probs = e / e.sum(dim=-1, keepdim=True)
    m = _mult(dropout, "attn", 4 * i + 1, probs)
    if m is not None:
        probs = probs * m
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, S, D)
    y = linear(ctx, P[pre + "attention.output.dense.weight"], P[pre + "attention.output.dense.bias"])
    m = _mult(dropout, "hidden", 4 * i + 2, y)
    if m is not None:
        y = y * m
    a = layer_norm(y + h, P[pre + "attention.output.LayerNorm.weight"], P[pre + "attention.output.LayerNorm.bias"])
    u = gelu(linear(a, P[pre + "intermediate.dense.weight"], P[pre + "intermediate.dense.bias"]))
    y = linear(u, P[pre + "output.dense.weight"], P[pre + "output.dense.bias"])
    m = _mult(dropout, "hidden", 4 * i + 3, y)
    if m is not None:
        y = y * m
    return layer_norm(y + a, P[pre + "output.LayerNorm.weight"], P[pre + "output.LayerNorm.bias"])


def forward(P, txt, mask, segment, img_tokens, cfg, indices=None, dropout=None):
    """``MultimodalBertClf.forward*`` from the pooled image tokens (B, N, d_img) onward.
    ``indices``: LongTensor over the full sequence (None = all).  ``dropout`` (train mode):
    ``dict(hidden=, attn=, img=, seed=)`` -- ImageBertEmbeddings.dropout (src/mmbt.py:82) on the
    image side of the embedded sequence, BertEmbeddings' on the text side (site 0, element counter
    over the SELECTED rows), then the per-layer sites of ``bert_layer``."""
    pre = "enc.txt_embeddings."
    B = txt.shape[0]
    n_img = img_tokens.shape[1]
    dtype = img_tokens.dtype
    full_mask = torch.cat([torch.ones(B, n_img + 2, dtype=torch.long), mask], dim=1)
    emb = torch.cat([image_bert_embeddings(P, img_tokens, cfg["cls_id"], cfg["sep_id"]),
                     bert_text_embeddings(P, txt, segment)], dim=1)
    emb = layer_norm(emb, P[pre + "LayerNorm.weight"], P[pre + "LayerNorm.bias"])
    pos = torch.arange(emb.shape[1])
    if indices is not None:
        emb, full_mask, pos = emb[:, indices], full_mask[:, indices], pos[indices]
    if dropout is not None:
        m_img, m_txt = _mult(dropout, "img", 0, emb), _mult(dropout, "hidden", 0, emb)
        one = torch.ones_like(emb)
        side = (pos < n_img + 2)[None, :, None]
        emb = emb * torch.where(side, m_img if m_img is not None else one, m_txt if m_txt is not None else one)
    ext = (1.0 - full_mask[:, None, None, :].to(dtype)) * -10000.0
    h = emb
    for i in range(cfg["n_layers"]):
        h = bert_layer(P, i, h, ext, cfg["n_head"], dropout)
    pooled = torch.tanh(linear(h[:, 0], P["enc.pooler.dense.weight"], P["enc.pooler.dense.bias"]))
    return linear(pooled, P["clf.weight"], P["clf.bias"])


def cross_entropy(logits, y):
    z = logits - logits.max(dim=-1, keepdim=True)[0]
    lse = torch.log(torch.exp(z).sum(-1))
    return (lse - z.gather(1, y[:, None])[:, 0]).mean()


def loss_and_grads(P, txt, mask, segment, img_tokens, y, cfg, indices=None, dropout=None):
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in P.items()}
    img_tokens = img_tokens.detach().clone().requires_grad_(True)
    logits = forward(P, txt, mask, segment, img_tokens, cfg, indices, dropout)
    loss = cross_entropy(logits, y)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()
             if v.is_floating_point()}
    # nn.Embedding(padding_idx=0): the gradient of word row 0 is dropped (bert_restated.BertEmbeddings)
    grads["enc.txt_embeddings.word_embeddings.weight"][0] = 0
    return logits.detach(), loss.detach(), grads, img_tokens.grad


def bertadam_step(p, g, m, v, step, *, lr, warmup, t_total, weight_decay, b1=0.9, b2=0.999, e=1e-6,
                  max_grad_norm=1.0):
    """One BertAdam update of ONE tensor (oracle/bert_restated.py::BertAdam.step)."""
    if max_grad_norm > 0:
        coef = max_grad_norm / (g.norm() + 1e-6)
        if coef < 1:
            g = g * coef
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    upd = m / (v.sqrt() + e)
    if weight_decay > 0:
        upd = upd + weight_decay * p
    x = step / t_total if t_total != -1 else None
    sched = 1.0 if x is None else (x / warmup if x < warmup else 1.0 - x)
    return p - lr * sched * upd, m, v
