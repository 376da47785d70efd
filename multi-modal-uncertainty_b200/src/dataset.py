"""Batch shaping for the hot path (reference ``src/dataset.py:30-101, 216-226``) and synthetic
stand-ins for the real-data loaders (which are out of scope: SURVEY.md section 2, row 6).

The shuffles use the HOST generator exactly like the reference (``torch.randperm`` on the CPU
default generator, same calls in the same order) so permutations and tiled labels are
bit-exact; only the resulting index tensors travel to the GPU.
"""
import torch
from torch.utils.data import Dataset


def data_forming_func_transformer(x, y, phase, model_type):
    """Reference src/dataset.py:30-54.  Vanilla: labels (B,) -> (B, 1); MultiHead: -> (B, 2);
    MIMO-shuffle-instance: image and text streams permuted independently (image perm drawn
    first), labels stacked per stream.  Eval phases pass through unchanged."""
    img, txt = x
    if phase == "train":
        if model_type == "Vanilla":
            y = y.unsqueeze(1).repeat(1, 1)
        elif model_type == "MultiHead":
            y = y.unsqueeze(1).repeat(1, 2)
        elif model_type == "MIMO-shuffle-instance":
            n = img.size(0)
            p_img = torch.randperm(n)
            p_txt = torch.randperm(n)
            # NB: the reference indexes `txt`/`y` with the second permutation after `img` has
            # already been permuted; the label pairing below is identical.
            img, txt = img[p_img], txt[p_txt]
            y = torch.stack([y[p_img], y[p_txt]], dim=1)
    return (img, txt), y


def data_forming_func(x, y, phase, model_type):
    """Reference src/dataset.py:56-101: four-view FashionMNIST variants."""
    b, m, c, h, w = x.shape
    train = phase == "train"
    if model_type == "single-model-weight-sharing":
        return x.reshape(-1, c, h, w), y.unsqueeze(1).repeat(1, m).reshape(-1)
    if not train:
        return x, y
    if model_type == "Vanilla":
        return x, y.unsqueeze(1).repeat(1, 1)
    if model_type == "MultiHead":
        return x, y.unsqueeze(1).repeat(1, m)
    if model_type == "MIMO-shuffle-view":
        return x[:, torch.randperm(m)], y.unsqueeze(1).repeat(1, m)
    if model_type in ("MIMO-shuffle-instance", "MIMO-shuffle-all"):
        views = 4 if model_type == "MIMO-shuffle-instance" else m
        perms = [torch.randperm(b) for _ in range(views)]
        x = torch.stack([x[p, i] for i, p in enumerate(perms)], dim=1)
        y = torch.stack([y[p] for p in perms], dim=1)
        if model_type == "MIMO-shuffle-all":
            vp = torch.randperm(x.size(1))
            x, y = x[:, vp], y[:, vp]
    return x, y


def quarter_views(x):
    """The four-view FashionMNIST input format (reference ``QuarterCrop`` + per-crop ``ToTensor``
    stack inside ``get_fmnist``, src/dataset.py:105-151) as a tensor op: ``(B, C, H, W)`` ->
    ``(B, 4, C, H/2, W/2)`` (or ``(C, H, W)`` -> ``(4, C, H/2, W/2)`` for one sample), views in the
    reference's order upper left, upper right, lower left, lower right.  Works on host or device
    tensors (slicing only), so a loader can keep whole 28 x 28 images in pinned memory and split
    them after the copy."""
    if x.dim() not in (3, 4) or x.shape[-1] % 2 or x.shape[-2] % 2:
        raise ValueError("quarter_views expects (C, H, W) or (B, C, H, W) with even H and W")
    hh, hw = x.shape[-2] // 2, x.shape[-1] // 2
    return torch.stack([x[..., :hh, :hw], x[..., :hh, hw:], x[..., hh:, :hw], x[..., hh:, hw:]], dim=x.dim() - 3)


def collate_fn_flava(batch):
    """Reference src/dataset.py:216-226: zero-pad ragged (l_i, 768) embeddings to the batch max."""
    def pad(seqs):
        out = seqs[0].new_zeros(len(seqs), max(s.shape[0] for s in seqs), seqs[0].shape[1])
        for i, s in enumerate(seqs):
            out[i, : s.shape[0]] = s
        return out
    imgs = pad([b[0] for b in batch])
    txts = pad([b[1] for b in batch])
    labels = torch.tensor([int(b[2]) for b in batch])
    return (imgs, txts), labels


def collate_fn(batch):
    """Reference src/dataset.py:420-438, the MMBT batch format: rows ``(tokens (l_i,), segment (l_i,),
    image, label (1,))`` as ``JsonlDataset.__getitem__`` returns them -> ``((text, segment, mask,
    image), target)`` with the three token-axis tensors zero-padded to the batch maximum (int64)
    and ``mask`` = 1 on real tokens.  (``Model_`` calls ``model(*x)`` on this tuple, so ``segment``
    lands in the model's ``mask`` argument and vice versa -- src/framework.py:172-176; the two are
    equal, ones on real tokens, so nothing changes.)"""
    lens = [int(row[0].shape[0]) for row in batch]
    text = torch.nn.utils.rnn.pad_sequence([row[0].long() for row in batch], batch_first=True)
    segment = torch.nn.utils.rnn.pad_sequence([row[1].long() for row in batch], batch_first=True)
    mask = (torch.arange(max(lens)).unsqueeze(0) < torch.tensor(lens).unsqueeze(1)).long()
    img = torch.stack([row[2] for row in batch])
    tgt = torch.cat([row[3] for row in batch]).long()
    return (text, segment, mask, img), tgt


class SyntheticFlavaDataset(Dataset):
    """Synthetic stand-in for ``FlavaEncodedDataset`` (reference src/dataset.py:196-213): the same
    item protocol (image_embeddings (l_img, d), text_embeddings (l_txt_i, d), LongTensor([label]))
    with seeded N(0, 1) features and optionally ragged text lengths."""

    def __init__(self, n, l_img=197, l_txt=40, dim=768, num_classes=101, seed=42, ragged=False):
        self.n, self.l_img, self.l_txt, self.dim, self.C = n, l_img, l_txt, dim, num_classes
        self.seed, self.ragged = seed, ragged
        self.num_classes = num_classes

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        g = torch.Generator().manual_seed(self.seed * 1000003 + idx)
        l_txt = self.l_txt
        if self.ragged:
            l_txt = int(torch.randint(max(1, self.l_txt // 2), self.l_txt + 1, (1,), generator=g))
        img = torch.randn(self.l_img, self.dim, generator=g)
        txt = torch.randn(l_txt, self.dim, generator=g)
        label = torch.randint(0, self.C, (1,), generator=g)
        return img, txt, label


def get_synthetic_flava(batch_size, n_train, n_val, n_test, seed=42, shuffle=True, **kw):
    """Loaders shaped like reference ``get_dataset`` (src/dataset.py:287-321): DataLoader with
    ``collate_fn_flava``, workers 0, generator seeded with ``seed``."""
    from torch.utils.data import DataLoader
    g = torch.Generator().manual_seed(seed)
    mk = lambda n, s, sh: DataLoader(SyntheticFlavaDataset(n, seed=s, **kw), batch_size=batch_size,
                                     shuffle=sh, collate_fn=collate_fn_flava, generator=g,
                                     drop_last=False)
    return mk(n_train, seed, shuffle), mk(n_val, seed + 1, False), mk(n_test, seed + 2, False)


# ------------------------------------------------------------------ packed embedding store
def pack_flava_encodings(samples, out_dir):
    """Write an iterable of ``(image_embeddings (l_img_i, d), text_embeddings (l_txt_i, d), label)``
    -- what ``data/encoding_with_flava.py:38-41`` saves as one ``.img`` + one ``.text`` file per
    sample -- as ONE packed store: ``image.f32`` / ``text.f32`` (rows back to back),
    ``image_offsets.npy`` / ``text_offsets.npy`` (int64, N+1) and ``labels.npy``.  Replaces two
    ``torch.load`` calls per sample (reference src/dataset.py:206-213) by two memory-mapped slices."""
    import os
    import numpy as np
    os.makedirs(out_dir, exist_ok=True)
    offs = {"image": [0], "text": [0]}
    labels, dim = [], None
    with open(os.path.join(out_dir, "image.f32"), "wb") as fi, \
            open(os.path.join(out_dir, "text.f32"), "wb") as ft:
        for img, txt, label in samples:
            img, txt = img.reshape(-1, img.shape[-1]), txt.reshape(-1, txt.shape[-1])
            dim = dim or img.shape[-1]
            assert img.shape[-1] == dim and txt.shape[-1] == dim
            fi.write(img.to(torch.float32).contiguous().numpy().tobytes())
            ft.write(txt.to(torch.float32).contiguous().numpy().tobytes())
            offs["image"].append(offs["image"][-1] + img.shape[0])
            offs["text"].append(offs["text"][-1] + txt.shape[0])
            labels.append(int(label))
    np.save(os.path.join(out_dir, "image_offsets.npy"), np.asarray(offs["image"], dtype=np.int64))
    np.save(os.path.join(out_dir, "text_offsets.npy"), np.asarray(offs["text"], dtype=np.int64))
    np.save(os.path.join(out_dir, "labels.npy"), np.asarray(labels, dtype=np.int64))
    np.save(os.path.join(out_dir, "dim.npy"), np.asarray([dim], dtype=np.int64))


class PackedFlavaDataset(Dataset):
    """Reader of :func:`pack_flava_encodings` stores with the item protocol of the reference's
    ``FlavaEncodedDataset`` (src/dataset.py:196-213): ``(image (l_img, d), text (l_txt_i, d),
    LongTensor([label]))``, served as views of two memory maps."""

    def __init__(self, store_dir):
        import os
        import numpy as np
        self.dim = int(np.load(os.path.join(store_dir, "dim.npy"))[0])
        self.off = {k: np.load(os.path.join(store_dir, k + "_offsets.npy")) for k in ("image", "text")}
        self.labels = np.load(os.path.join(store_dir, "labels.npy"))
        self.mm = {k: np.memmap(os.path.join(store_dir, k + ".f32"), dtype=np.float32, mode="r",
                                shape=(int(self.off[k][-1]), self.dim)) for k in ("image", "text")}

    def __len__(self):
        return len(self.labels)

    def rows(self, key, idx):
        return self.mm[key][int(self.off[key][idx]):int(self.off[key][idx + 1])]

    def __getitem__(self, idx):
        import numpy as np
        return (torch.from_numpy(np.array(self.rows("image", idx))),
                torch.from_numpy(np.array(self.rows("text", idx))),
                torch.LongTensor([int(self.labels[idx])]))


class RaggedCollator:
    """``collate_fn`` for :class:`PackedFlavaDataset` index batches: instead of padding on the
    host (reference ``collate_fn_flava``, src/dataset.py:216-226) it concatenates the samples'
    rows into one pinned buffer per modality and lets the GPU scatter them into the zero-padded
    (B, max_len, d) batch (``mmu_ragged_pad``) -- the host never writes a padding byte and the
    H2D copy moves only real tokens.  Use with ``DataLoader(range(len(ds)), collate_fn=...)``."""

    def __init__(self, dataset, device, pin=True):
        self.ds, self.device, self.pin = dataset, torch.device(device), pin

    def _modality(self, key, idxs):
        import numpy as np
        lens = [int(self.ds.off[key][i + 1] - self.ds.off[key][i]) for i in idxs]
        host = torch.empty(sum(lens), self.ds.dim, dtype=torch.float32, pin_memory=self.pin)
        o = 0
        for i, n in zip(idxs, lens):
            host[o:o + n] = torch.from_numpy(np.asarray(self.ds.rows(key, i)))
            o += n
        offsets = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32)
        if self.pin:
            offsets = offsets.pin_memory()
        from ._backend import ops
        return ops.ragged_pad(host.to(self.device, non_blocking=True),
                              offsets.to(self.device, non_blocking=True), max(lens))

    def __call__(self, idxs):
        idxs = [int(i) for i in idxs]
        labels = torch.tensor([int(self.ds.labels[i]) for i in idxs])
        return (self._modality("image", idxs), self._modality("text", idxs)), labels.to(self.device)


class DevicePrefetcher:
    """Iterates a loader of HOST batches ``((img, txt), y)`` and yields them on ``device``:
    batch i+1 is copied host->device on a side stream (from pinned memory: a true async DMA)
    while batch i is being consumed, so the copy overlaps the step's kernels instead of
    preceding them.  Replaces the per-step blocking ``.to(device)`` of the reference loop
    (src/framework.py:277-279); batches it yields pass through ``Model_.train_step`` /
    ``eval_step`` unchanged (``.to(device)`` on a resident tensor is a no-op).

    The device side is two persistent slots (no allocation in steady state; a slot is
    re-filled only after an event recorded behind the consumer's last use of it), so a yielded
    batch is valid until the consumer asks for the batch after the next one."""

    SLOTS = 2

    def __init__(self, loader, device, pin=True):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.stream = torch.cuda.Stream(device=self.device)
        self._bufs = [dict() for _ in range(self.SLOTS)]
        # event behind the consumer's last use of each slot; kept across __iter__ calls so that a
        # second pass over the loader never refills a slot the previous pass's last steps (still
        # queued on the GPU) are reading
        self._released = [None] * self.SLOTS

    def _copy(self, obj, slot, path):
        if obj is None:
            return None
        if isinstance(obj, (tuple, list)):
            return type(obj)(self._copy(o, slot, path + (i,)) for i, o in enumerate(obj))
        if self.pin and not obj.is_pinned():
            obj = obj.pin_memory()
        buf = self._bufs[slot].get(path)
        if buf is None or buf.shape != obj.shape or buf.dtype != obj.dtype:
            buf = torch.empty(obj.shape, dtype=obj.dtype, device=self.device)
            self._bufs[slot][path] = buf
            # the caching allocator may hand out a block that kernels ALREADY QUEUED on the
            # consumer's stream still use (freed by the host, not yet by the GPU): reuse is only
            # ordered on the allocating stream, so the copy stream must get behind that work
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            buf.copy_(obj, non_blocking=True)
        return buf

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        it = iter(self.loader)
        released = self._released
        state = {"k": 0}

        def fetch():
            try:
                host = next(it)
            except StopIteration:
                return None
            slot = state["k"] % self.SLOTS
            state["k"] += 1
            if released[slot] is not None:
                self.stream.wait_event(released[slot])
            dev = self._copy(host, slot, ())
            ready = torch.cuda.Event()
            ready.record(self.stream)
            return dev, ready, slot

        nxt = fetch()
        while nxt is not None:
            batch, ready, slot = nxt
            cur = torch.cuda.current_stream(self.device)
            nxt = fetch()            # batch i+1 starts copying now (into the other slot)
            cur.wait_event(ready)    # batch i has landed before the step touches it
            try:
                yield batch
            finally:        # also when the consumer breaks out of the loop
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
                released[slot] = ev
