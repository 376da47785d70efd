"""Robustness sweeps over token subsets / modality masks (reference
``eval_transformer_robustness.py:37-52, 95-137`` and ``eval_robustness.py:82-121``).

Index sets are drawn on the HOST from the same generators, in the same order, as the reference
(NumPy global state for n, torch CPU generator for the permutations) so masks are bit-exact;
the token gather, the forward passes and the scoring run on the GPU.  Instead of dumping
``(S, 43, K, C)`` logits to ``.npy`` for notebooks, metrics accumulate on device per variant
(one small all-reduce per sweep); the ``.npy``-shaped array is still returned on request.
"""
import numpy as np
import torch

from .metrics import PosthocMeter, UncertaintyMeter


def input_sampling(l_img, l_txt, type="image"):
    """Reference eval_transformer_robustness.py:37-52."""
    assert type in ("image", "text")
    l = l_img if type == "image" else l_txt
    n = int(np.random.randint(0, l + 1, size=1)[0])
    n_img, n_txt = (n, l - n) if type == "image" else (l - n, n)
    idx_img = torch.sort(torch.randperm(l_img)[:n_img]).values
    idx_txt = torch.sort(torch.randperm(l_txt)[:n_txt]).values
    return idx_img, idx_txt


def robustness_variants(l_img, l_txt, n_repeats=20):
    """[(idx_img | None, idx_txt | None)] in the reference's order (:103-121): full, image only,
    text only, n_repeats image-controlled draws, n_repeats text-controlled draws."""
    fi, ft = torch.arange(l_img), torch.arange(l_txt)
    out = [(fi, ft), (fi, None), (None, ft)]
    for type in ("image", "text"):
        for _ in range(n_repeats):
            ii, it = input_sampling(l_img, l_txt, type)
            out.append((ii if len(ii) else None, it if len(it) else None))
    return out


def mask_level_variant(l_img, l_txt, type, level, levels=10):
    """North-star config 3 grid: n_k = round(k l / (levels-1)) tokens of the
    controlled modality, l - n_k of the other; permutations from torch's CPU generator."""
    l = l_img if type == "image" else l_txt
    n = int(round(level * l / (levels - 1)))
    n_img, n_txt = (n, l - n) if type == "image" else (l - n, n)
    n_img, n_txt = min(n_img, l_img), min(n_txt, l_txt)
    idx_img = torch.sort(torch.randperm(l_img)[:n_img]).values
    idx_txt = torch.sort(torch.randperm(l_txt)[:n_txt]).values
    return (idx_img if n_img else None, idx_txt if n_txt else None)


def modality_dropout_mask(batch_size, p_drop, mode="random", scores=None, generator=None):
    """(B, 2) int32 keep mask over (image, text); guided / random modality dropout.  The
    reference only names these (configs/training_guided.gin:10-18); the definition is this
    repo's (see oracle/shaping.py) -- parity unpinned."""
    u = torch.rand(batch_size, generator=generator)
    pick = torch.rand(batch_size, generator=generator)
    keep = torch.ones(batch_size, 2, dtype=torch.int32)
    if mode == "random":
        which = (pick >= 0.5).to(torch.int64)
    elif mode == "guided":
        which = (scores[:, 1] > scores[:, 0]).to(torch.int64)
    else:
        raise ValueError(mode)
    rows = torch.nonzero(u < p_drop).flatten()
    keep[rows, which[rows]] = 0
    return keep


def modality_dropout_mask_device(batch_size, p_drop, mode, device, score_img=None, score_txt=None,
                                 generator=None):
    """The same keep mask as :func:`modality_dropout_mask`, built ON THE DEVICE: the two uniform
    vectors come from the host generator (same draws, same order), are uploaded from pinned
    memory, and ``mmu_modality_keep_mask`` combines them with device-resident guided scores
    (``score_img`` / ``score_txt``: fp32 (B,) views of equal stride, e.g. the confidence column of
    the uncertainty epilogue's per-sample output for the image-only / text-only variants of the
    batch) -- a guided training step never waits for a device->host read."""
    from ._backend import ops
    staged = torch.empty(2, batch_size, dtype=torch.float32, pin_memory=True)
    torch.rand(batch_size, generator=generator, out=staged[0])
    torch.rand(batch_size, generator=generator, out=staged[1])
    ur = staged.to(device, non_blocking=True)
    return ops.modality_keep_mask(ur[0], ur[1], p_drop, mode, score_img, score_txt)


@torch.no_grad()
def forward_variant(model, img, txt, variant, ref_bug_compat=False):
    """Run one variant.  ``ref_bug_compat`` reproduces reference line 119 (text slot indexed
    from ``img``); the default follows the evident intent."""
    ii, it = variant
    src_txt = img if ref_bug_compat else txt
    x = (img if ii is not None else None, src_txt if it is not None else None)
    full_i = ii is None or (len(ii) == img.shape[1] and bool((ii == torch.arange(len(ii))).all()))
    full_t = it is None or (len(it) == src_txt.shape[1] and bool((it == torch.arange(len(it))).all()))
    idx = (None if full_i else ii, None if full_t else it)
    return model(x, token_indices=idx)


@torch.no_grad()
def forward_variants(model, img, txt, variants, ref_bug_compat=False):
    """Logits (V, B, E, C) of every variant of one batch.  Uses the model's packed-variant pass
    when it has one; ``ref_bug_compat`` as in :func:`forward_variant`."""
    src_txt = img if ref_bug_compat else txt
    if hasattr(model, "forward_variants"):
        return model.forward_variants((img, src_txt), variants)
    return torch.stack([forward_variant(model, img, txt, v, ref_bug_compat) for v in variants])


@torch.no_grad()
def run_transformer_robustness(model, batches, device, n_repeats=20, ref_bug_compat=False,
                               collect=True, variants_fn=None, posthoc=False):
    """Sweep every batch through the variant schedule.  Returns (preds (S, V, K, C) numpy or
    None, labels numpy, [per-variant metric dicts]); with ``posthoc=True`` a fourth element holds
    the notebooks' scores (Pearson r of experimental vs control delta-p per modality, accuracy
    table) accumulated on device, so ``collect=False`` runs need no (S, 43, K, C) dump at all."""
    model.eval()
    meters, preds, labels = None, [], []
    # binary heads also get the notebooks' AUROC table (hatefulmeme_robustness.py:22-41)
    scorer = PosthocMeter(device, n_repeats, auc=model.num_classes == 2) if posthoc else None
    for (img, txt), y in batches:
        img, txt, y = img.to(device), txt.to(device), y.to(device).reshape(-1)
        if variants_fn is None:
            variants = robustness_variants(img.shape[1], txt.shape[1], n_repeats)
        else:
            variants = variants_fn(img.shape[1], txt.shape[1])
        if meters is None:
            meters = [UncertaintyMeter(device, model.num_classes, model.out_dim) for _ in variants]
        # all variants of the batch in one packed pass (model.forward_variants), then one fused
        # epilogue launch per variant into that variant's accumulator
        all_logits = forward_variants(model, img, txt, variants, ref_bug_compat)
        for logits, meter in zip(all_logits, meters):
            meter.update(logits, y)
        if scorer is not None:
            scorer.update(all_logits, y)
        if collect:
            preds.append(all_logits.transpose(0, 1).cpu())
        labels.append(y.cpu())
    for m in meters or []:
        m.all_reduce()
    P = torch.cat(preds).numpy() if collect and preds else None
    out = (P, torch.cat(labels).numpy(), [m.compute() for m in meters or []])
    if scorer is not None:
        scorer.all_reduce()
        out += (scorer.compute(),)
    return out


def view_sweep_inputs(x, i, model_type=None):
    """Input of sweep step i (reference ``eval_robustness.py:88-112``) for a ``(B, m, c, h, w)``
    batch of views.  Multi-head models: view i zero-filled, every other view kept (:92-97).
    ``single-model-weight-sharing``: view i REMOVED and the remaining m - 1 views laid out as the
    ``(B * (m - 1), c, h, w)`` batch the shared single-view model takes (:99-110; what
    ``data_forming_func(x_, y, 'eval', model_type)`` returns)."""
    if model_type == "single-model-weight-sharing":
        b, m, c, h, w = x.shape
        keep = [j for j in range(m) if j != i]
        return x[:, keep].reshape(b * (m - 1), c, h, w)
    x_ = x.clone()
    x_[:, i] = 0
    return x_


@torch.no_grad()
def run_view_robustness(model, batches, device, n_views=4, collect=True, model_type=None, metrics=True):
    """FashionMNIST four-view sweep (reference ``eval_robustness.py:82-126``): for view i, zero-fill
    view i of every sample (``model_type="single-model-weight-sharing"``: remove it and run the
    shared single-view model on the other three, logits regrouped to ``(B, 3, C)``, :99-114),
    forward, collect the logits.  Returns (outputs ``(n_views, S, E, C)`` numpy or None, labels
    numpy, [per-view metric dicts]) -- the arrays the reference saves as
    ``{ckpt}_predictions_robustness.npy`` / ``{ckpt}_labels.npy`` (with weight sharing the labels
    are the ones ``data_forming_func`` returns, each repeated m - 1 times, as the reference
    stores them); the metrics accumulate on device per left-out view (``metrics=False``: none,
    as in the reference)."""
    model.eval()
    shared = model_type == "single-model-weight-sharing"
    meters = None
    outs, labels = [[] for _ in range(n_views)], []
    for x, y in batches:
        x, y = x.to(device), y.to(device).reshape(-1)
        b, m = x.shape[0], x.shape[1]
        for i in range(n_views):
            logits = model(view_sweep_inputs(x, i, model_type))
            if shared:
                logits = logits.reshape(b, m - 1, logits.shape[-1])
            if metrics:
                if meters is None:
                    meters = [UncertaintyMeter(device, logits.shape[-1], logits.shape[1]) for _ in range(n_views)]
                meters[i].update(logits, y)
            if collect:
                outs[i].append(logits.cpu())
        labels.append((y.unsqueeze(1).repeat(1, m - 1).reshape(-1) if shared else y).cpu())
    for mt in meters or []:
        mt.all_reduce()
    P = torch.stack([torch.cat(o) for o in outs]).numpy() if collect and labels else None
    return P, torch.cat(labels).numpy(), [mt.compute() for mt in meters or []]


@torch.no_grad()
def run_predictions(model, batches, device, model_type=None, metrics=True):
    """Plain FashionMNIST prediction dump (reference ``eval_prediction_saving.py:77-104``): eval-mode
    forward of every batch, logits ``(S, M, C)`` and labels ``(S,)`` -- the arrays the reference saves
    as ``{ckpt}_predictions.npy`` / ``{ckpt}_labels.npy``.  With
    ``model_type="single-model-weight-sharing"`` the shared single-view model sees the
    ``(B * m, c, h, w)`` batch ``data_forming_func`` builds and its logits are regrouped to
    ``(B, m, C)`` (:84-95).  Returns (outputs numpy, labels numpy, metric dict | None); the
    uncertainty / calibration metrics of the dump accumulate on the device (``metrics=False``:
    none, as in the reference)."""
    from .dataset import data_forming_func
    model.eval()
    shared = model_type == "single-model-weight-sharing"
    meter, outs, labels = None, [], []
    for x, y in batches:
        b, m = x.shape[0], x.shape[1]
        y = y.reshape(-1)
        x_, _ = data_forming_func(x, y, "eval", model_type)
        logits = model(x_.to(device))
        if shared:
            logits = logits.reshape(b, m, logits.shape[-1])
        if metrics:
            if meter is None:
                meter = UncertaintyMeter(device, logits.shape[-1], logits.shape[1])
            meter.update(logits, y.to(device))
        outs.append(logits.cpu())
        labels.append(y.cpu())
    if meter is not None:
        meter.all_reduce()
    return torch.cat(outs).numpy(), torch.cat(labels).numpy(), (meter.compute() if meter is not None else None)


@torch.no_grad()
def run_mmbt_robustness(model, generator, n_repeats=20, device=None, posthoc=None):
    """The MMBT sweep of reference ``eval_mmbt_robustness.py:76-96``: per batch the full forward,
    ``forward_img_only``, ``forward_txt_only`` and ``n_repeats`` ``forward_control`` draws for
    "image" then "text" -> ``(S, 3 + 2 n_repeats, C)`` logits and ``(S,)`` labels (the arrays the
    reference saves as ``robustness_{ckpt}_predictions_{phase}.npy`` / ``..._labels_...``).

    The reference re-runs the ResNet-152 image encoder inside every one of the 43 forwards; in
    eval mode its output does not depend on the variant, so the pooled image tokens are computed
    ONCE per batch and every variant is an index list over the same embedded sequence
    (``MultimodalBertClf.forward_indices``).  The ``forward_control`` draws consume the host
    ``torch.randperm`` stream exactly like the reference (src/mmbt.py:198-201)."""
    model.eval()
    device = device if device is not None else next(model.parameters()).device
    preds, labels = [], []
    for x, y in generator:
        txt, mask, segment, img = (t.to(device, non_blocking=True) for t in x)
        tokens = model._tokens(img)
        n_img, s_txt = model._n_img, txt.shape[1]
        total = s_txt + n_img + 2
        # index lists of the 43 variants in the reference's order; the forward_control draws come
        # from the host RNG in the order the reference's calls would make them
        lists = [list(range(total)), list(range(n_img + 2)), [0] + list(range(n_img + 2, total))]
        for type in ("image", "text"):
            num = n_img + 1 if type == "image" else s_txt
            for _ in range(n_repeats):
                lists.append(model.control_indices(total, num))
        if hasattr(model, "forward_index_lists"):
            # equally long variants share ONE pass (the 21 five-token variants -- img_only and the
            # "image" controls -- would otherwise be 21 launch-bound forwards)
            outs = [None] * len(lists)
            by_len = {}
            for v, il in enumerate(lists):
                by_len.setdefault(len(il), []).append(v)
            for length, vs in by_len.items():
                if len(vs) > 1 and length * len(vs) <= total:
                    packed = model.forward_index_lists(txt, mask, segment, tokens, [lists[v] for v in vs])
                    for k, v in enumerate(vs):
                        outs[v] = packed[k]
                else:
                    for v in vs:
                        outs[v] = model.forward_indices(txt, mask, segment, tokens,
                                                        None if length == total else lists[v])
        else:
            outs = [model.forward_indices(txt, mask, segment, tokens, il) for il in lists]
        y_hat = torch.stack(outs, dim=1)  # (B, V, C)
        if posthoc is not None:
            posthoc.update(y_hat.transpose(0, 1).unsqueeze(2).contiguous(), y.to(device))
        preds.append(y_hat.cpu())
        labels.append(y.cpu())
    return torch.cat(preds).numpy(), torch.cat(labels).numpy()
