"""Callback protocol of the training framework (reference ``src/callbacks.py:16-356``): the same
hook names and setter methods, so user callbacks written for the reference plug in unchanged.
Host-side bookkeeping only -- nothing here is on the GPU path."""
import sys
import timeit

import numpy as np

from .utils import save_weights


class Callback:
    """Hooks: on_train_begin/end, on_epoch_begin/end, on_batch_begin/end,
    on_forward_begin(batch, data), on_backward_end(batch), on_val_batch_end."""

    def set_meta_data(self, meta_data): self.meta_data = meta_data
    def set_save_path(self, save_path): self.save_path = save_path
    def set_optimizer(self, optimizer): self.optimizer = optimizer
    def set_model_pytoune(self, model_pytoune): self.model_pytoune = model_pytoune
    def set_params(self, params): self.params = params
    def set_dataloader(self, data): self.data = data

    def set_model(self, model, ignore=True):
        if not ignore:
            self.model = model

    def get_meta_data(self): return self.meta_data
    def get_dataloader(self): return self.data
    def get_optimizer(self): return self.optimizer
    def get_params(self): return self.params
    def get_model(self): return self.model
    def get_save_path(self): return self.save_path

    def on_train_begin(self, logs): pass
    def on_train_end(self, logs): pass
    def on_epoch_begin(self, epoch, logs): pass
    def on_epoch_end(self, epoch, logs): pass
    def on_batch_begin(self, batch, logs): pass
    def on_batch_end(self, batch, logs): pass
    def on_forward_begin(self, batch, data): pass
    def on_backward_end(self, batch): pass
    def on_val_batch_end(self, batch, logs): pass


_HOOKS = ("on_train_begin", "on_train_end", "on_epoch_begin", "on_epoch_end", "on_batch_begin",
          "on_batch_end", "on_forward_begin", "on_backward_end", "on_val_batch_end")
_SETTERS = ("set_params", "set_model", "set_model_pytoune", "set_optimizer", "set_save_path")


class CallbackList:
    """Fan-out container (reference src/callbacks.py:16-80)."""

    def __init__(self, callbacks=None):
        self.callbacks = list(callbacks or [])

    def append(self, callback):
        self.callbacks.append(callback)

    def __iter__(self):
        return iter(self.callbacks)

    def __getattr__(self, name):
        if name in _HOOKS or name in _SETTERS:
            def fan_out(*args, **kwargs):
                for cb in self.callbacks:
                    getattr(cb, name)(*args, **kwargs)
            return fan_out
        raise AttributeError(name)


class LambdaCallback(Callback):
    """Reference src/callbacks.py:154-186: build a callback from keyword lambdas."""

    def __init__(self, **hooks):
        for name, fn in hooks.items():
            if name not in _HOOKS:
                raise ValueError(f"unknown hook {name}")
            if fn is not None:
                setattr(self, name, fn)


class ModelCheckpoint(Callback):
    """Save {'model','optimizer'} when the monitored value improves (or every `period` epochs);
    reference src/callbacks.py:188-254 (which breaks on NumPy >= 2 through ``np.Inf``).  With
    ``save_best_only=False`` the reference's ``save_weights`` call sits under ``if self.verbose > 0``
    (:252-254, an indentation slip: nothing is written at verbose 0); here every period saves.
    Pickles without its model / optimizer references, like the reference (:217-226)."""

    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, mode="auto",
                 period=1):
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.period = save_best_only, period
        self.epochs_since_last_save = 0
        if mode not in ("min", "max"):
            mode = "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"
        self.monitor_op = np.greater if mode == "max" else np.less
        self.best = -np.inf if mode == "max" else np.inf

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("model", None)
        state.pop("optimizer", None)
        return state

    def __setstate__(self, newstate):
        for k in ("model", "optimizer"):   # keep the live references of the unpickling object, if any
            if k in self.__dict__:
                newstate[k] = self.__dict__[k]
        self.__dict__.update(newstate)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        self.epochs_since_last_save = 0
        if not self.save_best_only:
            save_weights(self.model, self.optimizer, self.filepath)
            return
        current = logs.get(self.monitor)
        if current is None:
            return
        if self.monitor_op(current, self.best):
            if self.verbose:
                print(f"Epoch {epoch:05d}: {self.monitor} improved {self.best:.5f} -> {current:.5f}")
            self.best = current
            save_weights(self.model, self.optimizer, self.filepath)


class ProgressionCallback(Callback):
    """Console progress (reference src/callbacks.py:256-316), reduced to one line per epoch and an
    ETA per `every` batches; silent unless ``verbose``."""

    def __init__(self, other_metrics=(), verbose=True, every=50):
        self.other_metrics = list(other_metrics)
        self.verbose, self.every = verbose, every

    def on_train_begin(self, logs):
        self.epochs = self.params.get("epochs")
        self.steps = self.params.get("steps")

    def on_epoch_begin(self, epoch, logs):
        self.epoch, self.t0, self.seen, self.loss_sum = epoch, timeit.default_timer(), 0, 0.0

    def on_batch_end(self, batch, logs):
        self.seen += logs["size"]
        self.loss_sum += logs["loss"] * logs["size"]
        if self.verbose and self.steps and batch % self.every == 0:
            dt = timeit.default_timer() - self.t0
            eta = dt / batch * (self.steps - batch)
            sys.stdout.write(f"\rEpoch {self.epoch}/{self.epochs} step {batch}/{self.steps} "
                             f"ETA {eta:.0f}s loss: {self.loss_sum / max(self.seen, 1):.6f}")
            sys.stdout.flush()

    def on_epoch_end(self, epoch, logs):
        if self.verbose:
            keys = ["loss", "acc", "val_loss", "val_acc", "test_acc"] + self.other_metrics
            msg = " ".join(f"{k}: {logs[k]:.6f}" for k in keys if k in logs)
            print(f"\rEpoch {epoch}/{self.epochs} {logs.get('time', 0.0):.2f}s {msg}")


class ValidationProgressionCallback(Callback):
    """Reference src/callbacks.py:318-356 (per-batch validation progress); quiet by default."""

    def __init__(self, phase, metrics_names, steps=None, verbose=False):
        self.phase, self.metrics_names, self.steps, self.verbose = phase, metrics_names, steps, verbose

    def on_batch_end(self, batch, logs):
        if self.verbose and self.steps:
            sys.stdout.write(f"\r{self.phase} {batch}/{self.steps}")
            sys.stdout.flush()
