"""Oracle, "eager incumbent" form (TEST / BASELINE INFRASTRUCTURE -- never imported by the product).

The same arithmetic as ``oracle/fusion.py`` but expressed with the FUSED ATen operators the
reference's ``nn.Module`` tree dispatches to (``F.linear``, ``F.layer_norm``,
``F.multi_head_attention_forward`` -> scaled-dot-product attention, ``F.cross_entropy``), on
whatever device / dtype its inputs live on.  Two uses only:

* ``bench.py``'s ``incumbent`` leg times it on ``cuda`` -- "PyTorch eager on B200 running the
  reference modules" (SURVEY 2.1), i.e. cuBLAS + ATen, the real bar to beat;
* ``tests/test_oracle_golden.py`` pins it to ``oracle.fusion`` (and through it to the reference
  goldens), so the incumbent that is timed provably computes the reference's function.

Reference: src/model.py:188-212 (block), :258-304 (forward, loss); train.py:196-202 (AdamW).
"""
import torch
import torch.nn.functional as F


def block(x, P, pre, n_head):
    """ResidualAttentionBlock.forward (src/model.py:205-212), attention over axis 0."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), P[pre + "ln_1.weight"], P[pre + "ln_1.bias"], 1e-5)
    a, _ = F.multi_head_attention_forward(
        h, h, h, D, n_head, P[pre + "attn.in_proj_weight"], P[pre + "attn.in_proj_bias"], None, None,
        False, 0.0, P[pre + "attn.out_proj.weight"], P[pre + "attn.out_proj.bias"],
        training=False, need_weights=False)
    x = x + a
    h = F.layer_norm(x, (D,), P[pre + "ln_2.weight"], P[pre + "ln_2.bias"], 1e-5)
    z = F.linear(h, P[pre + "mlp.c_fc.weight"], P[pre + "mlp.c_fc.bias"])
    u = z * torch.sigmoid(1.702 * z)
    return x + F.linear(u, P[pre + "mlp.c_proj.weight"], P[pre + "mlp.c_proj.bias"])


def forward(P, x, n_head, n_layers, n_out, avg_pool=False):
    """FlavaFusionTransfomer.forward (src/model.py:258-291); a ``None`` modality is dropped."""
    img, txt = x
    parts, l_img = [], 0
    if img is not None:
        parts.append(F.linear(img, P["image_to_mm_projection.weight"], P["image_to_mm_projection.bias"]))
        l_img = img.shape[1]
    if txt is not None:
        parts.append(F.linear(txt, P["text_to_mm_projection.weight"], P["text_to_mm_projection.bias"]))
    mm = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
    D = mm.shape[-1]
    mm = F.layer_norm(mm, (D,), P["ln_pre.weight"], P["ln_pre.bias"], 1e-5)
    for i in range(n_layers):
        mm = block(mm, P, f"mm_encoder.resblocks.{i}.", n_head)
    out = F.layer_norm(mm, (D,), P["ln_post.weight"], P["ln_post.bias"], 1e-5)
    if avg_pool:
        feats = [out[:, :l_img].mean(1), out[:, l_img:].mean(1)]
    else:
        feats = [out[:, i] for i in range(n_out)]
    return torch.stack([F.linear(f, P[f"output_layers.{i}.weight"], P[f"output_layers.{i}.bias"])
                        for i, f in enumerate(feats)], dim=1)


def compute_loss(y_hat, y, eval=False):
    """src/model.py:293-304."""
    if not eval:
        return F.cross_entropy(y_hat.reshape(-1, y_hat.shape[2]), y.reshape(-1))
    return F.cross_entropy(y_hat.mean(1), y.reshape(-1))


class EagerTrainer:
    """Train step + sweep of the benchmark, the way the reference runs them: autograd over the
    forward above, ``torch.optim.AdamW`` (fused=False: the reference's torch 1.12 had no fused
    kernel; ``fused=True`` is offered as the stronger incumbent), one forward per sweep level."""

    def __init__(self, P, n_head, n_layers, n_out, lr, wd, fused_optimizer=False, autocast=None):
        self.P = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        self.n_head, self.n_layers, self.n_out = n_head, n_layers, n_out
        self.autocast = autocast
        kw = dict(fused=True) if fused_optimizer else {}
        self.opt = torch.optim.AdamW(list(self.P.values()), lr=lr, betas=(0.9, 0.98), eps=1e-9,
                                     weight_decay=wd, **kw)

    def _ctx(self, device_type):
        if self.autocast is None:
            import contextlib
            return contextlib.nullcontext()
        return torch.autocast(device_type, dtype=self.autocast)

    def train_step(self, x, yt):
        self.opt.zero_grad(set_to_none=True)
        with self._ctx(yt.device.type):
            logits = forward(self.P, x, self.n_head, self.n_layers, self.n_out)
            loss = compute_loss(logits.float(), yt)
        loss.backward()
        self.opt.step()
        pred = logits.reshape(-1, logits.shape[2]).argmax(1)
        return loss.detach(), (pred == yt.reshape(-1)).float().mean() * 100

    @torch.no_grad()
    def sweep(self, x, variants):
        img, txt = x
        outs = []
        with self._ctx(img.device.type):
            for ii, it in variants:
                s_img = img[:, ii.to(img.device)] if ii is not None else None
                s_txt = txt[:, it.to(txt.device)] if it is not None else None
                outs.append(forward(self.P, (s_img, s_txt), self.n_head, self.n_layers, self.n_out))
        return torch.stack(outs).float()
