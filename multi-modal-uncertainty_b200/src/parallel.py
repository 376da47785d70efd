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
    gradient all-reduce of each finished bucket on a communication stream; ``optimizer.step()``
    waits for them and applies the averaged gradient.

    Measured on 8 x B200 (``tools/ddp_probe.py`` -> ``profiles/r02_ddp_attribution_8gpu.json``):
    the four per-stage all-reduces of the 91 MB fp32 gradient take 0.37 ms back to back and are
    hidden almost completely by this overlap -- 0.06 ms of a 5.4 ms training step stays exposed.
    What separates the 8-GPU step from the 1-GPU step is NOT communication: with the gradient
    exchange switched off the same step is 3.3 % slower when all eight GPUs of the box are busy
    (5.37 vs 5.20 ms), and the collective-free evaluation sweep 2.5 % (8.92 vs 8.70 ms).

    Knobs that were measured and are OFF by default because they lost: ``nccl_max_ctas`` (a
    communicator of its own with ``ncclConfig_t.maxCTAs`` capped: the collectives slow down more
    than the backward speeds up, +0.2 .. +0.8 ms), ``reserve_sms`` (GEMM grids sized to the SMs
    NCCL leaves free, ``mmu_set_gemm_sm_limit``: neutral), ``merge_stages`` (fewer, larger buckets:
    the last one is exposed, +0.1 .. +0.5 ms).
    """

    def __init__(self, model, optimizer, group=None, overlap=True, nccl_max_ctas=0, reserve_sms=0,
                 merge_stages=1, record_events=False):
        self.model, self.optimizer, self.overlap = model, optimizer, overlap
        self.rank, self.world = world(group)
        self.group = group
        cuda = model._flat.is_cuda
        if self.world > 1 and cuda and group is None and nccl_max_ctas and dist.get_backend() == "nccl":
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = int(nccl_max_ctas)
            opts.config.min_ctas = 1
            self.group = dist.new_group(ranks=list(range(self.world)), pg_options=opts)
        self.reserve_sms = int(reserve_sms or 0) if cuda else 0
        if not (self.world > 1 and overlap):
            self.reserve_sms = 0
        self.ranges = model.stage_ranges()
        self.buckets = self._make_buckets(self.ranges, max(1, int(merge_stages)))
        self._works = []
        self._comm = torch.cuda.Stream() if (overlap and cuda) else None
        self.record_events = record_events
        self.events = []          # [(bucket, bytes, start_event, end_event)] of the last backward
        broadcast_flat(model._flat, 0, self.group)
        optimizer.grad_scale = 1.0 / self.world
        model._ddp = self

    @staticmethod
    def _make_buckets(ranges, merge):
        """[(last_stage, begin, end)]: the all-reduce of elements [begin, end) may start once the
        backward stage ``last_stage`` has finished.  Stages are visited in order 0..n-1 and (by the
        flat layout) finish contiguous ranges, so merged buckets stay contiguous; the last (stem)
        stage always joins the bucket before it."""
        groups, cur = [], []
        for st in range(len(ranges)):
            cur.append(st)
            if len(cur) >= merge and st < len(ranges) - 2:
                groups.append(cur)
                cur = []
        if cur:
            groups.append(cur)
        out = []
        for g in groups:
            b = min(ranges[st][0] for st in g)
            e = max(ranges[st][1] for st in g)
            out.append((g[-1], b, e))
        return out

    def _launch(self, model, bucket):
        last, b, e = bucket
        if self._comm is None:
            self._works += all_reduce_ranges(model._flat_grad, [(b, e)], self.group, True)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            if self.record_events:
                # a synchronous-op collective makes THIS stream wait for NCCL's own stream, so the
                # closing event sees the end of the all-reduce kernel (the host never blocks)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                all_reduce_ranges(model._flat_grad, [(b, e)], self.group, False)
                t1.record()
                self.events.append((last, (e - b) * 4, t0, t1))
            else:
                self._works += all_reduce_ranges(model._flat_grad, [(b, e)], self.group, True)

    def backward(self, model, cfg, inp, ws, dlogits):
        n = len(self.ranges)
        if self.world == 1:
            model.backward_stages(cfg, inp, ws, dlogits, 0, n)
            return
        self.events = []
        if self.reserve_sms:
            _lib.lib.mmu_set_gemm_sm_limit(max(2, (_sm_count(model._flat.device) - self.reserve_sms) // 2 * 2))
        try:
            done = 0
            for bucket in self.buckets:
                model.backward_stages(cfg, inp, ws, dlogits, done, bucket[0] + 1)
                done = bucket[0] + 1
                self._launch(model, bucket)
        finally:
            if self.reserve_sms:
                _lib.lib.mmu_set_gemm_sm_limit(0)

    def wait(self):
        for w in self._works:
            w.wait()
        self._works = []
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)

    def bucket_times(self):
        """[(last_stage, bytes, ms)] of the all-reduces of the last backward (``record_events``)."""
        torch.cuda.synchronize()
        return [(st, nbytes, t0.elapsed_time(t1)) for st, nbytes, t0, t1 in self.events]


def _sm_count(device):
    return torch.cuda.get_device_properties(device).multi_processor_count


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
