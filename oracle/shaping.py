"""Oracle: batch shaping, padding and robustness index sampling (TEST INFRASTRUCTURE).

All randomness in the reference comes from the HOST generators (``torch.randperm`` on the CPU
default generator, ``np.random.randint`` on NumPy's global state), so the restatement draws
from the same generators in the same order; masks and permutations are then bit-exact by
construction and the tests compare integer tensors for equality.
"""
import numpy as np
import torch


# ------------------------------------------------------------------ training-time shaping
def data_forming_func_transformer(x, y, phase, model_type):
    """src/dataset.py:30-54.  Label tiling for multi-head training; for MIMO the image and text
    streams are shuffled independently (two ``randperm`` draws, image first)."""
    img, txt = x
    if phase == "train":
        if model_type == "Vanilla":
            y = y.unsqueeze(1).repeat(1, 1)
        elif model_type == "MultiHead":
            y = y.unsqueeze(1).repeat(1, 2)
        elif model_type == "MIMO-shuffle-instance":
            perm_i = torch.randperm(img.size(0))
            img, y_img = img[perm_i], y[perm_i]
            perm_t = torch.randperm(img.size(0))
            txt, y_txt = txt[perm_t], y[perm_t]
            y = torch.stack([y_img, y_txt], dim=1)
    return (img, txt), y


def data_forming_func(x, y, phase, model_type):
    """src/dataset.py:56-101 (four-view FashionMNIST variants)."""
    b, m, c, h, w = x.shape
    train = phase == "train"
    if model_type == "Vanilla" and train:
        y = y.unsqueeze(1).repeat(1, 1)
    elif model_type == "single-model-weight-sharing":
        y = y.unsqueeze(1).repeat(1, m).reshape(-1)
        x = x.reshape(-1, c, h, w)
    elif model_type == "MultiHead" and train:
        y = y.unsqueeze(1).repeat(1, m)
    elif model_type == "MIMO-shuffle-instance" and train:
        xs, ys = [], []
        for i in range(4):
            perm = torch.randperm(x.size(0))
            xs.append(x[perm, i])
            ys.append(y[perm])
        x, y = torch.stack(xs, dim=1), torch.stack(ys, dim=1)
    elif model_type == "MIMO-shuffle-view" and train:
        x = x[:, torch.randperm(x.size(1))]
        y = y.unsqueeze(1).repeat(1, m)
    elif model_type == "MIMO-shuffle-all" and train:
        xs, ys = [], []
        for i in range(m):
            perm = torch.randperm(x.size(0))
            xs.append(x[perm, i])
            ys.append(y[perm])
        x, y = torch.stack(xs, dim=1), torch.stack(ys, dim=1)
        view_perm = torch.randperm(x.size(1))
        x, y = x[:, view_perm], y[:, view_perm]
    return x, y


def collate_fn_flava(batch):
    """src/dataset.py:216-226: zero-pad ragged (l_i, D) embeddings to the batch maximum."""
    def pad(seqs):
        lmax = max(s.shape[0] for s in seqs)
        out = torch.zeros(len(seqs), lmax, seqs[0].shape[1], dtype=seqs[0].dtype)
        for i, s in enumerate(seqs):
            out[i, : s.shape[0]] = s
        return out

    imgs = pad([b[0] for b in batch])
    txts = pad([b[1] for b in batch])
    labels = torch.tensor([int(b[2]) for b in batch])
    return (imgs, txts), labels


def collate_fn(batch):
    """src/dataset.py:420-438 (MMBT): rows (tokens (l_i,), segment (l_i,), image, label (1,)) ->
    ((text, segment, mask, image), target); token / segment / mask rows zero-padded to the batch
    maximum, all int64."""
    lens = [int(row[0].shape[0]) for row in batch]
    n, lmax = len(batch), max(lens)
    text = torch.zeros(n, lmax, dtype=torch.int64)
    segment = torch.zeros(n, lmax, dtype=torch.int64)
    mask = torch.zeros(n, lmax, dtype=torch.int64)
    for i, (row, l) in enumerate(zip(batch, lens)):
        for t in range(l):
            text[i, t] = int(row[0][t])
            segment[i, t] = int(row[1][t])
            mask[i, t] = 1
    img = torch.stack([row[2] for row in batch])
    tgt = torch.cat([row[3] for row in batch]).long()
    return (text, segment, mask, img), tgt


# --------------------------------------------------------------------- robustness sweeps
def input_sampling(l_img, l_txt, type="image"):
    """eval_transformer_robustness.py:37-52.  n ~ U{0..l} from NumPy's global state, then two
    sorted ``randperm`` prefixes (image first) from torch's CPU generator."""
    assert type in ("image", "text")
    l = l_img if type == "image" else l_txt
    n = int(np.random.randint(0, l + 1, size=1)[0])
    n_img = n if type == "image" else l - n
    n_txt = n if type == "text" else l - n
    idx_img = torch.sort(torch.randperm(l_img)[:n_img]).values
    idx_txt = torch.sort(torch.randperm(l_txt)[:n_txt]).values
    return idx_img, idx_txt


def robustness_variants(l_img, l_txt, n_repeats=20):
    """The per-batch variant schedule of eval_transformer_robustness.py:99-121, as a list of
    (idx_img | None, idx_txt | None); ``None`` means "modality absent".  Order: full, image
    only, text only, then n_repeats "image" draws, then n_repeats "text" draws (3 + 2 n)."""
    full_i, full_t = torch.arange(l_img), torch.arange(l_txt)
    variants = [(full_i, full_t), (full_i, None), (None, full_t)]
    for type in ("image", "text"):
        for _ in range(n_repeats):
            ii, it = input_sampling(l_img, l_txt, type)
            variants.append((ii if len(ii) > 0 else None, it if len(it) > 0 else None))
    return variants


def apply_variant(img, txt, variant, ref_bug_compat=False):
    """Gather the token subsets of one variant.  ``ref_bug_compat`` reproduces line 119 of the
    reference script, which indexes ``img`` with the TEXT indices."""
    ii, it = variant
    s_img = img[:, ii, :] if ii is not None else None
    src_txt = img if ref_bug_compat else txt
    s_txt = src_txt[:, it, :] if it is not None else None
    return s_img, s_txt


def mask_level_variant(l_img, l_txt, type, level, levels=10):
    """North-star config 3: a deterministic grid over ``input_sampling``'s n ~ U{0..l}
    (SURVEY 8d.3): keep n_k = round(k l / (levels-1)) tokens of the controlled modality and
    l - n_k of the other, index sets from torch's CPU generator (image first)."""
    l = l_img if type == "image" else l_txt
    n = int(round(level * l / (levels - 1)))
    n_img = n if type == "image" else l - n
    n_txt = n if type == "text" else l - n
    n_img, n_txt = min(n_img, l_img), min(n_txt, l_txt)
    idx_img = torch.sort(torch.randperm(l_img)[:n_img]).values
    idx_txt = torch.sort(torch.randperm(l_txt)[:n_txt]).values
    return (idx_img if n_img > 0 else None, idx_txt if n_txt > 0 else None)


def leave_one_view_out(x, i):
    """eval_robustness.py:92-97: zero-fill view i of a (B, 4, 1, 14, 14) batch."""
    out = torch.zeros_like(x)
    for j in range(x.shape[1]):
        if j != i:
            out[:, j] = x[:, j]
    return out


def drop_one_view(x, i):
    """eval_robustness.py:99-108 (``single-model-weight-sharing``): view i REMOVED -- the remaining
    m - 1 views of a (B, m, c, h, w) batch, in order, as a (B, m - 1, c, h, w) tensor."""
    b, m, c, h, w = x.shape
    out = torch.zeros(b, m - 1, c, h, w, dtype=x.dtype)
    k = 0
    for j in range(m):
        if j != i:
            out[:, k] = x[:, j]
            k += 1
    return out


def quarter_views(img):
    """src/dataset.py:105-151: ``QuarterCrop((28, 28))`` crops (top, left) = (0, 0), (0, 14), (14, 0),
    (14, 14) -- upper left, upper right, lower left, lower right -- of a (..., 28, 28) image, each
    its own view: (..., 28, 28) -> (4, ..., 14, 14) stacked like the reference's per-crop ToTensor."""
    hh, hw = img.shape[-2] // 2, img.shape[-1] // 2
    crops = []
    for top, left in ((0, 0), (0, hw), (hh, 0), (hh, hw)):
        crops.append(img[..., top:top + hh, left:left + hw])
    return torch.stack(crops, dim=0)


# ------------------------------------------------- modality dropout (north-star extension)
def modality_dropout_mask(batch_size, p_drop, mode="random", scores=None, generator=None):
    """Per-sample keep mask (B, 2) int32 over (image, text).  NO REFERENCE IMPLEMENTATION
    (only stale gin names, configs/training_guided.gin:10-18) -- parity unpinned; this is the
    repo's DEFINITION, written sample by sample (the product computes it vectorised / on device):

      draw u_0..u_{B-1}, then r_0..r_{B-1}, uniform in [0, 1) from ``generator`` (two vector draws);
      sample b keeps both modalities unless u_b < p_drop; if it drops one,
        "random": image when r_b < 0.5, else text;
        "guided": the modality whose score is currently HIGHER (scores[b] = (image, text);
                  ties -> image), i.e. the one the model leans on.
    A dropped modality is zero-filled (``apply_keep_mask``), as eval_robustness.py:92-97
    zero-fills views."""
    if mode not in ("random", "guided"):
        raise ValueError(mode)
    u = torch.rand(batch_size, generator=generator).tolist()
    r = torch.rand(batch_size, generator=generator).tolist()
    rows = []
    for b in range(batch_size):
        keep_img, keep_txt = 1, 1
        if u[b] < p_drop:
            if mode == "random":
                drop_text = not (r[b] < 0.5)
            else:
                assert scores is not None and tuple(scores.shape) == (batch_size, 2)
                drop_text = bool(float(scores[b, 1]) > float(scores[b, 0]))
            if drop_text:
                keep_txt = 0
            else:
                keep_img = 0
        rows.append([keep_img, keep_txt])
    return torch.tensor(rows, dtype=torch.int32).reshape(batch_size, 2)


def apply_keep_mask(img, txt, keep):
    """Zero-fill the dropped modality of each sample (keep (B, 2) over (image, text))."""
    k = keep.to(img.dtype)
    return img * k[:, 0].view(-1, 1, 1), txt * k[:, 1].view(-1, 1, 1)
