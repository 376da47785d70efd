"""Data parallelism over one 8xB200 box: one process per GPU, batch sharding, NCCL over NVLink.

The reference has no distributed code at all; this adds exactly the two
collectives the path needs:

* training  -- sum-all-reduce of the flat fp32 gradient buffer, issued per backward STAGE
  (heads, each block in reverse order, stem) on a side stream so it overlaps the rest of the
  backward pass; the 1/world_size factor is folded into the fused AdamW kernel;
* evaluation -- one sum-all-reduce of the metric / histogram accumulators at the end of a
  sweep.  The integer bins are order independent, so the merged histogram is bit-exact.

There is no exchange step inside the model: attention spans the rank-local mini-batch, exactly
as ``DistributedDataParallel`` around the reference would behave.  The helpers work on any
backend (the CPU test-suite drives them with gloo, world_size 2).
"""
import torch
import torch.distributed as dist

from ._backend import _lib


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous [begin, end) slice of n units owned by `rank` (remainder to the low ranks)."""
    base, rem = divmod(n, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_reduce_accum(accum, group=None):
    """Sum an ``mmu_metric_accum`` (int64 view, see ``ops.new_accum``) across ranks in place."""
    if world(group)[1] == 1:
        return accum
    ints = accum[:_lib.ACC_INT_WORDS]
    dbls = accum[_lib.ACC_INT_WORDS:].view(torch.float64)
    dist.all_reduce(ints, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(dbls, op=dist.ReduceOp.SUM, group=group)
    return accum


def all_gather_rows(t, group=None):
    """Concatenation over ranks (rank order) of per-rank tensors that differ in their number of
    rows: what a rank statistic over the whole evaluation set needs (AUROC / Kendall tau do not
    decompose over sample shards, unlike the histogram accumulators).  Rows are padded to the
    longest shard for one ``all_gather``; 4 bytes per sample-variant cross NVLink."""
    _, nranks = world(group)
    if nranks == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(nranks)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def all_reduce_ranges(flat, ranges, group=None, async_op=False):
    """Sum-all-reduce the element ranges [(b, e), ...] of a flat buffer; returns work handles."""
    works = []
    for b, e in ranges:
        w = dist.all_reduce(flat[b:e], op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def broadcast_flat(flat, src=0, group=None):
    if world(group)[1] > 1:
        dist.broadcast(flat, src=src, group=group)


class DataParallel:
    """Attach to a model + FusedAdamW pair:  ``ddp = DataParallel(model, optimizer)``.

    After that ``loss.backward()`` runs the engine backward stage by stage and launches the
    gradient all-reduce of each finished stage on a communication stream; ``optimizer.step()``
    waits for them and applies the averaged gradient."""

    def __init__(self, model, optimizer, group=None, overlap=True):
        self.model, self.optimizer, self.group, self.overlap = model, optimizer, group, overlap
        self.rank, self.world = world(group)
        self.ranges = model.stage_ranges()
        self._works = []
        self._comm = torch.cuda.Stream() if (overlap and model._flat.is_cuda) else None
        broadcast_flat(model._flat, 0, group)
        optimizer.grad_scale = 1.0 / self.world
        model._ddp = self

    def backward(self, model, cfg, inp, ws, dlogits):
        n = len(self.ranges)
        if self.world == 1:
            model.backward_stages(cfg, inp, ws, dlogits, 0, n)
            return
        for st in range(n):
            model.backward_stages(cfg, inp, ws, dlogits, st, st + 1)
            b, e = self.ranges[st]
            if self._comm is None:
                self._works += all_reduce_ranges(model._flat_grad, [(b, e)], self.group, True)
                continue
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ev)
                self._works += all_reduce_ranges(model._flat_grad, [(b, e)], self.group, True)

    def wait(self):
        for w in self._works:
            w.wait()
        self._works = []
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)


class FlatGradSync:
    """Data parallelism for models made of several flat-buffer owners (the MMBT path: BERT trunk +
    image encoder) with the fused ``BertAdam``: parameters (and BatchNorm running statistics) are
    broadcast from rank 0 once; ``optimizer.step()`` then sum-all-reduces every owner's flat
    gradient buffer (one collective per owner) and the 1/world factor is folded into the update
    kernel.  BatchNorm batch statistics stay rank-local, as under ``DistributedDataParallel``
    without SyncBN.  ``attach(model, optimizer)`` wires it up."""

    def __init__(self, owners, group=None, overlap=False):
        self.owners, self.group, self.overlap = list(owners), group, overlap
        self.rank, self.world = world(group)
        self._pending = {}
        self._comm = None
        for o in self.owners:
            broadcast_flat(o._flat, 0, group)
            stats = getattr(o, "_stats", None)
            if stats is not None:
                broadcast_flat(stats, 0, group)
            if hasattr(o, "invalidate_shadow"):
                o.invalidate_shadow()
            o._flat_sync = self if overlap else None

    def launch(self, owner):
        """Called by an owner at the end of its backward (``overlap=True`` only -- one backward per
        optimizer step, i.e. no gradient accumulation): its gradient all-reduce starts on a side
        stream and overlaps whatever backward work is still queued (the image encoder's backward
        runs after the trunk's)."""
        if self.world == 1 or id(owner) in self._pending:
            return
        if self._comm is None and owner._flat_grad.is_cuda:
            self._comm = torch.cuda.Stream()
        if self._comm is None:
            self._pending[id(owner)] = dist.all_reduce(owner._flat_grad, group=self.group, async_op=True)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            self._pending[id(owner)] = dist.all_reduce(owner._flat_grad, group=self.group, async_op=True)

    def all_reduce_grads(self):
        if self.world > 1:
            for o in self.owners:
                w = self._pending.pop(id(o), None)
                if w is not None:
                    w.wait()
                else:
                    dist.all_reduce(o._flat_grad, group=self.group)
            if self._comm is not None:
                torch.cuda.current_stream().wait_stream(self._comm)

    @classmethod
    def attach(cls, optimizer, group=None, overlap=False):
        sync = cls(optimizer._owners, group, overlap)
        optimizer._flat_sync = sync
        optimizer.grad_scale = 1.0 / sync.world
        return sync
