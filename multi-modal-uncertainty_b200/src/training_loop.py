"""Default callbacks, history.csv and checkpoint loading (reference ``src/training_loop.py``):
same artefacts -- ``history.csv``, ``model_epoch_{n}.pt``, ``model_last_epoch.pt``,
``model_best_val.pt`` -- and the same strict state-dict reload."""
import logging
import os
from functools import partial

import numpy as np
import pandas as pd
import torch

from .callbacks import LambdaCallback, ModelCheckpoint
from .utils import save_weights

logger = logging.getLogger(__name__)
_CSV_TYPES = (int, float, complex, np.integer, np.floating, str)


def _append_to_history(epoch, logs, H):
    for key, value in logs.items():
        H.setdefault(key, []).append(value)


def _save_history_csv(epoch, logs, save_path, H):
    logger.info("\t".join(f"{k}={v}" for k, v in logs.items() if isinstance(v, _CSV_TYPES)))
    keep = {k: v for k, v in H.items() if isinstance(v[-1], _CSV_TYPES)}
    pd.DataFrame(keep).to_csv(os.path.join(save_path, "history.csv"), index=False)


def _construct_default_callbacks(model, optimizer, H, save_path, checkpoint_monitor):
    """Reference src/training_loop.py:23-47."""
    def save_every_epoch(epoch, logs):
        save_weights(model, optimizer, os.path.join(save_path, f"model_epoch_{epoch}.pt"))
        save_weights(model, optimizer, os.path.join(save_path, "model_last_epoch.pt"))

    return [
        LambdaCallback(on_epoch_end=partial(_append_to_history, H=H)),
        LambdaCallback(on_epoch_end=partial(_save_history_csv, save_path=save_path, H=H)),
        ModelCheckpoint(monitor=checkpoint_monitor, save_best_only=True, mode="max",
                        filepath=os.path.join(save_path, "model_best_val.pt")),
        LambdaCallback(on_epoch_end=save_every_epoch),
    ]


def _load_pretrained_model(model, save_path):
    """Reference src/training_loop.py:72-77: strict reload of checkpoint['model']."""
    checkpoint = torch.load(save_path, map_location="cpu")
    state = model.state_dict()
    state.update(checkpoint["model"])
    model.load_state_dict(state, strict=True)
    logger.info("Done reloading!")
