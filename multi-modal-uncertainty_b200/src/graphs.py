"""CUDA-graph capture of a whole training step for the small-kernel configurations.

The FashionMNIST models (reference ``train_fashionmnist.py``: four 14x14 views, batch 256,
``torch.optim.SGD``) spend 110 MFLOP per sample in ~170 small kernels per step (SURVEY.md 8d,
8f.4).  ``GraphedTrainStep`` records the step body of ``Model_.train_step`` --
``optimizer.zero_grad()``, forward, ``compute_loss``, ``backward()``, ``optimizer.step()``, metrics
(reference ``src/framework.py:262-312``) -- once per batch shape into a ``torch.cuda.CUDAGraph``
and replays it with one launch per step (forward, loss and backward go through
``model.forward_backward``, which makes the same engine calls without the autograd engine); the
host only copies the batch into the graph's static input buffers.  The same kernels run in the
same order on the same addresses as in the eager step, so results agree with it to the rounding of
the split-K atomics (which also separates two eager runs).  Measured on ``MIMOResNet`` at batch
256: 1.79 -> 1.54 ms per step in tensor-core mode, 7.8 -> 7.5 ms in fp32 mode (the step is mostly
bound by its kernels, not by launches: DESIGN.md 4.5).

What is baked into a captured graph, and therefore part of its cache key: the batch shapes /
dtypes and every scalar hyper-parameter of the optimiser's ``param_groups`` (a scheduler that moves
the learning rate -- ``ReduceLROnPlateau`` in ``train_fashionmnist.py:118`` -- triggers one
re-capture; after ``MAX_GRAPHS`` distinct sets new ones run eagerly, so a per-batch scheduler
degrades to the eager step instead of capturing a graph per batch).  Optimisers that compute
per-step scalars on the HOST from a step counter (``FusedAdamW`` / ``BertAdam`` bias corrections
and warm-up, ``torch.optim.Adam`` without ``capturable=True``) cannot be replayed and are
rejected.  The very first step always runs eagerly: it creates the optimiser state (momentum
buffers) the captured step updates in place.
"""
import torch


def _flatten(x):
    if isinstance(x, (tuple, list)):
        return list(x), True
    return [x], False


class GraphedTrainStep:
    MAX_GRAPHS = 8  # captured graphs kept (each owns its static buffers and a memory pool)

    def __init__(self, trainer):
        self.trainer = trainer
        self.entries = {}
        self._warm = False
        self._warned = False
        if not hasattr(trainer.model, "forward_backward"):
            raise TypeError("the model must provide forward_backward(x, y) (the fusion / FashionMNIST "
                            "models do) to be replayed from a CUDA graph")
        opt = trainer.optimizer
        name = type(opt).__name__
        if name in ("FusedAdamW", "BertAdam"):
            raise ValueError(f"{name} derives per-step scalars from a host-side step counter; a "
                             "replayed graph would freeze them -- use the eager step")
        for g in opt.param_groups:
            if "capturable" in g and not g["capturable"] and name != "SGD":
                raise ValueError(f"{name} must be built with capturable=True to be replayed from a CUDA graph")

    def _key(self, xs, y):
        shapes = tuple(None if t is None else (tuple(t.shape), t.dtype) for t in xs)
        hyper = tuple(tuple(sorted((k, v) for k, v in g.items()
                                   if isinstance(v, (int, float, bool)) or v is None))
                      for g in self.trainer.optimizer.param_groups)
        return shapes, tuple(y.shape), y.dtype, hyper

    def _baked(self):
        """Every tensor whose ADDRESS a captured step bakes in: the model's flat parameter /
        gradient / statistics buffers, the bf16 shadow, the cached workspaces and the optimiser
        state.  An entry keeps strong references to them (a freed workspace can therefore never be
        recycled under a live graph) and is dropped as soon as one of the model- or
        optimiser-owned buffers is no longer the current one (``model.to()``, ``_rebind``,
        ``optimizer.load_state_dict`` ...)."""
        m, opt = self.trainer.model, self.trainer.optimizer
        owned = [getattr(m, n, None) for n in ("_flat", "_flat_grad", "_stats", "_shadow")]
        for st in opt.state.values():
            owned += [v for v in st.values() if torch.is_tensor(v)]
        owned = [t for t in owned if t is not None]
        ws = getattr(m, "_ws", None)
        scratch = list(ws.values()) if isinstance(ws, dict) else []
        return owned, scratch

    @staticmethod
    def _ident(tensors):
        return tuple((t.data_ptr(), t.numel(), t.dtype) for t in tensors)

    def _body(self, x, y):
        tr = self.trainer
        tr.optimizer.zero_grad()
        # forward + loss + backward through the model's autograd-free entry point: the autograd
        # engine's stream bookkeeping makes the capture depend on uncaptured work
        # (cudaErrorStreamCaptureIsolation), the engine calls themselves are plain launches
        y_pred, loss = tr.model.forward_backward(x, y)
        tr.optimizer.step()
        with torch.no_grad():
            mets = [m(y_pred, y, False, True) for m in tr.metrics]
        return loss, mets

    def _capture(self, xs, is_seq, y):
        dev = self.trainer.device
        if dev is None or torch.device(dev).type != "cuda":
            raise RuntimeError("CUDA-graph capture needs the trainer on a CUDA device: call Model_.to(device) first")
        sx = [None if t is None else torch.empty_like(t, device=dev) for t in xs]
        sy = torch.empty_like(y, device=dev)
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        try:
            with torch.cuda.graph(graph):
                loss, mets = self._body(sx if is_seq else sx[0], sy)
        except RuntimeError as e:  # a metric / model that reads back to the host under capture
            raise RuntimeError("the training step cannot be captured in a CUDA graph (a host "
                               f"read-back or allocation inside the step?): {e}") from e
        bad = [m for m in mets if not torch.is_tensor(m)]
        if bad:
            raise TypeError("metrics must return device tensors to be replayed from a CUDA graph")
        owned, scratch = self._baked()
        return graph, sx, sy, loss, mets, self._ident(owned), owned + scratch

    def step(self, x, y):
        """x, y: the shaped batch (host or device tensors).  Returns (loss, metrics) as 0-dim
        device tensors that the NEXT replay overwrites: read them before calling again."""
        xs, is_seq = _flatten(x)
        if not self._warm:  # first step: eager (creates the optimiser state)
            self._warm = True
            tr = self.trainer
            return self._body(tr.to_device(x), tr.to_device(y))
        key = self._key(xs, y)
        ent = self.entries.get(key)
        if ent is not None and ent[5] != self._ident(self._baked()[0]):
            del self.entries[key]      # a baked-in buffer was replaced: the graph is stale
            ent = None
        if ent is None:
            if len(self.entries) >= self.MAX_GRAPHS:
                # a scheduler that moves the learning rate every batch (or ever-changing batch
                # shapes) would capture a new graph per step: run such steps eagerly instead
                if not self._warned:
                    self._warned = True
                    import warnings
                    warnings.warn(f"more than {self.MAX_GRAPHS} distinct (shape, hyper-parameter) sets: "
                                  "further new ones run eagerly instead of being captured")
                tr = self.trainer
                return self._body(tr.to_device(x), tr.to_device(y))
            ent = self.entries[key] = self._capture(xs, is_seq, y)
        graph, sx, sy, loss, mets = ent[:5]
        for dst, src in zip(sx, xs):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        sy.copy_(y, non_blocking=True)
        graph.replay()
        # replays do not move tensor version counters: caches keyed on them are stale now
        inval = getattr(self.trainer.model, "invalidate_shadow", None)
        if inval is not None:
            inval()
        return loss, mets
