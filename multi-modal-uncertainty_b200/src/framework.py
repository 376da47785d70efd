"""Training / evaluation driver with the reference's ``Model_`` protocol
(reference ``src/framework.py:35-355``): same constructor, ``to``, ``train_loop`` and
``eval_loop`` signatures, same callback hooks, same epoch-log keys (``loss, acc, val_loss,
val_acc, val_auc, test_*, time, epoch``) and the same stopping rules (NaN loss; train accuracy
== 100 for ``patience`` epochs).

The per-batch body is the hot path: ``data_forming`` (host RNG) -> H2D -> ``model(x)`` ->
``compute_loss`` -> ``backward`` -> ``optimizer.step`` -> metrics -> ``scheduler.step``.  It is
exposed on its own as ``train_step`` / ``eval_step`` so it can be benchmarked through the public
API.  The ``mmbt`` branch drives the MMBT engine; the ``vilt`` branch keeps the reference's calling
convention for transformers' ViLT classifier (dict batches, ``outputs.loss`` / ``.logits``).
"""
import itertools
import math
import timeit

import numpy as np
import torch

from .callbacks import CallbackList, ProgressionCallback, ValidationProgressionCallback


def _step_stream(steps, generator):
    if steps is None:
        return zip(itertools.count(1), generator)

    def cycle():
        while True:
            yield from generator
    return zip(range(1, steps + 1), cycle())


class StepIterator:
    """Yields ``(step, batch)``; after the body fills ``step['loss'|'metrics'|'size']`` it keeps
    size-weighted running means and fires the batch callbacks (reference :35-95).

    Deferred read-back (``read_every`` > 1): when the body hands over the loss /
    metrics as 0-dim DEVICE tensors, the size-weighted sums are accumulated ON THE DEVICE and read
    back once every ``read_every`` steps and at the end of the epoch (one small D2H copy) instead
    of the reference's two queue-draining ``.item()`` reads per step (src/framework.py:305-312).
    The epoch means are the same numbers; a NaN loss is detected at the next read (``saw_nan``);
    per-batch logs carry the running mean of the last read."""

    _reserved = ("loss", "metrics", "number", "size")

    def __init__(self, generator, steps_per_epoch, callback, metrics_names, read_every=1):
        self.generator, self.steps_per_epoch = generator, steps_per_epoch
        self.callback, self.metrics_names = callback, metrics_names
        self.losses_sum, self.sizes_sum = 0.0, 0.0
        self.metrics_sum = np.zeros(len(metrics_names))
        self.extra_lists = {}
        self.read_every = max(1, int(read_every))
        self._dev, self._pending, self._pending_sizes = None, 0, 0.0
        self.saw_nan = False

    def _defer(self, loss, metrics, size):
        vec = torch.stack([loss.detach().reshape(()).float()] +
                          [torch.as_tensor(m, device=loss.device).reshape(()).float() for m in metrics])
        if self._dev is None:
            self._dev = torch.zeros_like(vec, dtype=torch.float64)
        self._dev.add_(vec, alpha=float(size))
        self._pending += 1
        self._pending_sizes += size
        if self._pending >= self.read_every:
            self.flush()

    def flush(self):
        """One device->host read of the pending size-weighted sums."""
        if self._pending:
            vals = self._dev.tolist()
            self._dev.zero_()
            self.losses_sum += vals[0]
            self.metrics_sum += np.asarray(vals[1:])
            self.sizes_sum += self._pending_sizes
            self.saw_nan = self.saw_nan or math.isnan(vals[0])
            self._pending, self._pending_sizes = 0, 0.0

    @property
    def loss(self):
        self.flush()
        return self.losses_sum / self.sizes_sum if self.sizes_sum else 0

    @property
    def metrics(self):
        self.flush()
        if not self.sizes_sum:
            return dict(zip(self.metrics_names, np.zeros(len(self.metrics_names))))
        return dict(zip(self.metrics_names, self.metrics_sum / self.sizes_sum))

    def __iter__(self):
        for number, data in _step_stream(self.steps_per_epoch, self.generator):
            t0 = timeit.default_timer()
            self.callback.on_batch_begin(number, {})
            self.callback.on_forward_begin(number, data)
            step = {"number": number}
            yield step, data
            size = step["size"]
            if torch.is_tensor(step["loss"]):      # deferred read-back: accumulate on the device
                self._defer(step["loss"], step["metrics"], size)
                mean_loss = self.losses_sum / self.sizes_sum if self.sizes_sum else float("nan")
                mean_mets = (self.metrics_sum / self.sizes_sum if self.sizes_sum
                             else np.full(len(self.metrics_names), np.nan))
                logs = {"batch": number, "size": size, "time": timeit.default_timer() - t0,
                        "batch_begin_time": t0, "loss": mean_loss,
                        **dict(zip(self.metrics_names, mean_mets)), "deferred": True}
            else:
                self.losses_sum += step["loss"] * size
                self.metrics_sum += step["metrics"] * size
                self.sizes_sum += size
                self.saw_nan = self.saw_nan or math.isnan(step["loss"])
                logs = {"batch": number, "size": size, "time": timeit.default_timer() - t0,
                        "batch_begin_time": t0, "loss": step["loss"],
                        **dict(zip(self.metrics_names, step["metrics"]))}
            for key, value in step.items():
                if key not in self._reserved:
                    self.extra_lists.setdefault(key, []).append(value)
            self.callback.on_batch_end(number, logs)


class Model_:
    def __init__(self, model, optimizer, scheduler, data_forming_func, *, metrics=[], verbose=True):
        self.model, self.optimizer, self.scheduler = model, optimizer, scheduler
        self.data_forming = data_forming_func
        self.metrics = metrics
        self.metrics_names = [m.__name__ for m in metrics]
        self.device = None
        self.verbose = verbose
        self.verbose_logs = {}

    # ---------------------------------------------------------------- plumbing
    def _compute_metrics(self, pred_y, y, eval, dummy_dim):
        return np.array([float(m(pred_y, y, eval, dummy_dim)) for m in self.metrics])

    def _transfer_optimizer_state_to_right_device(self):
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                for v in self.optimizer.state.get(p, {}).values():
                    if torch.is_tensor(v) and v.device != p.device:
                        v.data = v.data.to(p.device)

    def to(self, device):
        self.device = device
        self.model.to(device)
        for m in self.metrics:
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def to_device(self, x):
        if isinstance(x, (tuple, list)):
            return [None if t is None else t.to(self.device, non_blocking=True) for t in x]
        return x.to(self.device, non_blocking=True)

    # ---------------------------------------------------------------- ViLT branch
    # Reference src/framework.py:163-169, 263-272, 294-304 with train.py:164-182: the model is
    # transformers' ViltForImagesAndTextClassification (a third-party module with its own loss),
    # batches are dicts, ``outputs = model(**batch)`` carries ``.loss`` and ``.logits``, labels sit
    # in ``batch['labels']``, metrics see (B, C) logits (dummy_dim=False), no data_forming.  Every
    # model that follows that calling convention works; the arithmetic is the library's.
    def _vilt_batch(self, batch):
        return {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def train_step_vilt(self, batch, global_step, *, gradient_accumulation_steps=1, scheduler_step_on="epoch"):
        batch = self._vilt_batch(batch)
        y = batch["labels"]
        self.optimizer.zero_grad()
        outputs = self.model(**batch)
        y_pred, loss = outputs.logits, outputs.loss
        if gradient_accumulation_steps > 1:
            loss = loss / gradient_accumulation_steps
        loss.backward()
        if global_step % gradient_accumulation_steps == 0:
            self.optimizer.step()
            self.optimizer.zero_grad()
        with torch.no_grad():
            info = self._compute_metrics(y_pred, y, eval=False, dummy_dim=False)
        if scheduler_step_on == "batch" and self.scheduler is not None:
            self.scheduler.step()
        return loss.item(), info, len(y)

    @torch.no_grad()
    def eval_step_vilt(self, batch):
        batch = self._vilt_batch(batch)
        outputs = self.model(**batch)
        y = batch["labels"]
        info = self._compute_metrics(outputs.logits, y, eval=True, dummy_dim=False)
        return float(outputs.loss), info, len(y), outputs.logits, y

    # ---------------------------------------------------------------- hot path
    def train_step(self, x, y, scheduler_step_on="batch", keep_mask=None, sync=True, cuda_graph=False):
        """One optimisation step on a HOST (or prefetched device) batch; returns
        (loss: float, metrics: ndarray, size).  ``sync=False`` returns the loss and the metrics
        as 0-dim DEVICE tensors instead, so the caller chooses when to pay the device->host read
        (the reference reads ``loss.item()`` and every metric right here, src/framework.py:305-312,
        which drains the GPU queue twice per step).  ``cuda_graph=True`` replays the whole step
        body from a CUDA graph captured per batch shape (``graphs.GraphedTrainStep``: the
        FashionMNIST configuration; same kernels, same order as the eager step)."""
        x, y = self.data_forming(x, y, phase="train")
        if cuda_graph:
            if keep_mask is not None:
                raise ValueError("keep_mask batches are not replayed from a CUDA graph")
            if getattr(self, "_graphed", None) is None:
                from .graphs import GraphedTrainStep
                self._graphed = GraphedTrainStep(self)
            loss, info = self._graphed.step(x, y)
            if scheduler_step_on == "batch" and self.scheduler is not None:
                self.scheduler.step()
            if sync:
                return loss.item(), np.array([float(m) for m in info]), len(y)
            return loss, info, len(y)
        x, y = self.to_device(x), self.to_device(y)
        self.optimizer.zero_grad()
        y_pred = self.model(x) if keep_mask is None else self.model(x, keep_mask=keep_mask)
        loss = self.model.compute_loss(y_pred, y)
        loss.backward()
        self.optimizer.step()
        with torch.no_grad():
            if sync:
                info = self._compute_metrics(y_pred, y, eval=False, dummy_dim=True)
            else:
                info = [m(y_pred, y, False, True) for m in self.metrics]
        if scheduler_step_on == "batch" and self.scheduler is not None:
            self.scheduler.step()
        return (loss.item() if sync else loss.detach()), info, len(y)

    def train_step_mmbt(self, x, y, global_step, *, freeze_img=False, freeze_txt=False,
                        gradient_accumulation_steps=1, scheduler_step_on="epoch"):
        """The ``mmbt`` branch of the reference's step body (src/framework.py:277-304): freeze
        flags re-applied every step, ``model(*x)``, loss divided by the accumulation factor,
        optimizer step + zero_grad every ``gradient_accumulation_steps`` batches.  (The
        reference also zeroes the gradients at the START of every batch, :279 -- so its
        accumulation only ever sees the last micro-batch; reproduced as is.)"""
        x, y = self.data_forming(x, y, phase="train")
        x, y = self.to_device(x), self.to_device(y)
        self.optimizer.zero_grad()
        ie = getattr(self.model.enc, "img_encoder", None)
        if ie is not None:
            for p in ie.parameters():
                p.requires_grad = not freeze_img
        for p in self.model.enc.encoder.parameters():
            p.requires_grad = not freeze_txt
        y_pred = self.model(*x)
        loss = self.model.compute_loss(y_pred, y)
        if gradient_accumulation_steps > 1:
            loss = loss / gradient_accumulation_steps
        loss.backward()
        if global_step % gradient_accumulation_steps == 0:
            self.optimizer.step()
            self.optimizer.zero_grad()
        with torch.no_grad():
            info = self._compute_metrics(y_pred, y, eval=False, dummy_dim=False)
        if scheduler_step_on == "batch" and self.scheduler is not None:
            self.scheduler.step()
        return loss.item(), info, len(y)

    @torch.no_grad()
    def eval_step_mmbt(self, x, y):
        x, y = self.data_forming(x, y, phase="eval")
        x, y = self.to_device(x), self.to_device(y)
        outputs = self.model(*x)
        loss = self.model.compute_loss(outputs, y, eval=True)
        info = self._compute_metrics(outputs, y, eval=True, dummy_dim=False)
        return float(loss), info, len(y), outputs, y

    @torch.no_grad()
    def eval_step(self, x, y):
        x, y = self.data_forming(x, y, phase="eval")
        x, y = self.to_device(x), self.to_device(y)
        outputs = self.model(x)
        loss = self.model.compute_loss(outputs, y, eval=True)
        info = self._compute_metrics(outputs, y, eval=True, dummy_dim=True)
        return float(loss), info, len(y), outputs, y

    # ---------------------------------------------------------------- loops
    def eval_loop(self, generator, phase, *, steps=None, auc=False, mmbt=False, vilt=False):
        steps = len(generator) if steps is None else steps
        it = StepIterator(generator, steps,
                          ValidationProgressionCallback(phase=phase, steps=steps,
                                                        metrics_names=["loss"] + self.metrics_names),
                          self.metrics_names)
        self.model.eval()
        preds, labels = [], []
        for step, batch in it:
            if vilt:
                loss, info, size, outputs, y_dev = self.eval_step_vilt(batch)
            else:
                x, y = batch
                loss, info, size, outputs, y_dev = self.eval_step_mmbt(x, y) if mmbt else self.eval_step(x, y)
            step["size"], step["loss"], step["metrics"] = size, loss, info
            # (B, E, C) -> head-mean logits.  The reference applies the same .mean(1) to MMBT's
            # (B, C) logits (src/framework.py:191), which collapses the class axis and cannot feed
            # its own AUROC line (:198); the (B, C) logits are kept instead.
            preds.append(outputs if (mmbt or vilt) else outputs.mean(1))
            labels.append(y_dev)
        out = {f"{phase}_loss": it.loss,
               **{f"{phase}_{k}": v for k, v in it.extra_lists.items()},
               **{f"{phase}_{k}": v for k, v in it.metrics.items()}}
        if auc:
            # reference src/framework.py:195-198 (sklearn roc_auc_score on the host copy of every
            # logit): exact pair counting on the device instead, 32 bytes read back
            from . import metrics as _metrics
            out[f"{phase}_auc"] = _metrics.auroc(torch.cat(labels), torch.cat(preds)[:, 1].contiguous())
        return out

    def train_loop(self, train_generator, test_generator=None, valid_generator=None, *,
                   epochs=1000, steps_per_epoch=None, validation_steps=None, test_steps=None,
                   patience=10, callbacks=[], epoch_start=1, scheduler_step_on="epoch", auc=False,
                   mmbt=False, vilt=False, **kwargs):
        self._transfer_optimizer_state_to_right_device()
        cbs = CallbackList(callbacks)
        cbs.append(ProgressionCallback(verbose=self.verbose))
        cbs.set_params({"epochs": epochs, "steps": steps_per_epoch})
        cbs.set_model_pytoune(self)

        stop, stopped_epoch, perfect_epochs, global_step = False, 0, 0, 0
        cbs.on_train_begin({})
        for epoch in range(epoch_start, epochs + 1):
            cbs.on_epoch_begin(epoch, {})
            t0 = timeit.default_timer()
            # metrics_every=N (N > 1): loss / metrics stay on the device and are read back every N
            # steps and at the end of the epoch (StepIterator) -- same epoch means, no per-step sync
            read_every = 1 if (mmbt or vilt) else int(kwargs.get("metrics_every", 1))
            it = StepIterator(train_generator, steps_per_epoch, cbs, self.metrics_names, read_every)
            self.model.train(True)
            with torch.enable_grad():
                for step, batch in it:
                    x, y = (None, None) if vilt else batch
                    if vilt:
                        global_step += 1
                        loss, info, size = self.train_step_vilt(
                            batch, global_step, gradient_accumulation_steps=kwargs["gradient_accumulation_steps"],
                            scheduler_step_on=scheduler_step_on)
                    elif mmbt:
                        global_step += 1
                        loss, info, size = self.train_step_mmbt(
                            x, y, global_step, freeze_img=epoch < kwargs["freeze_img"],
                            freeze_txt=epoch < kwargs["freeze_txt"],
                            gradient_accumulation_steps=kwargs["gradient_accumulation_steps"],
                            scheduler_step_on=scheduler_step_on)
                    else:
                        loss, info, size = self.train_step(x, y, scheduler_step_on, sync=read_every == 1,
                                                           cuda_graph=bool(kwargs.get("cuda_graph", False)))
                    cbs.on_backward_end(step["number"])
                    step["size"], step["loss"], step["metrics"] = size, loss, info
            it.flush()
            if it.saw_nan:      # reference src/framework.py:319 (checked at every read-back)
                stop = True
            log = {"epoch": epoch, "loss": it.loss,
                   **{f"train_{k}": v for k, v in it.extra_lists.items()}, **it.metrics}
            if valid_generator is not None:
                log.update(self.eval_loop(valid_generator, "val", steps=validation_steps, auc=auc, mmbt=mmbt,
                                          vilt=vilt))
            if test_generator is not None:
                log.update(self.eval_loop(test_generator, "test", steps=test_steps, auc=auc, mmbt=mmbt, vilt=vilt))
            log["time"] = timeit.default_timer() - t0
            log["epoch_begin_time"] = t0
            if scheduler_step_on == "epoch" and self.scheduler is not None:
                self.scheduler.step(log[kwargs["scheduler_metric"]])
            cbs.on_epoch_end(epoch, log)
            if log.get("acc") == 100:
                perfect_epochs += 1
            if perfect_epochs >= patience:
                stopped_epoch, stop = epoch, True
            if stop:
                break
        cbs.on_train_end({})
        if stopped_epoch > 0:
            print("Epoch %05d: completed stopping" % stopped_epoch)
