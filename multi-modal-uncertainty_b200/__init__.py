"""B200-native hot path of wooginawunan/multi-modal-uncertainty (sm_100a only).

The directory name mirrors the reference repository (it contains hyphens), so import it with
``importlib.import_module("multi-modal-uncertainty_b200")`` or through the ``mmu_b200`` alias
module at the repository root.  ``src/`` mirrors the reference's ``src`` package (same class and
function names, constructor signatures, state-dict keys and call protocol) for the hot path;
everything numerical runs in ``libmmu_b200.so`` (hand-written CUDA, C ABI in
``include/mmu_b200.h``) -- importing this package fails if that library has not been built.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA library is missing)
from . import ops  # noqa: F401
from .src import (callbacks, dataset, framework, graphs, metrics, mmbt, model, optim, parallel,  # noqa: F401
                  robustness, training_loop, utils)
from .src.framework import Model_  # noqa: F401
from .src.metrics import acc  # noqa: F401
from .src.model import (FlavaFusionTransfomer, FlavaFusionTransfomerwithCLSToken,  # noqa: F401
                        MIMOTransfomer, model_configure)
from .src.resnet import MIMOResNet  # noqa: F401
from .src.mmbt import MultimodalBertClf  # noqa: F401
from .src.optim import BertAdam, FusedAdamW, get_cosine_schedule_with_warmup  # noqa: F401

__all__ = ["FlavaFusionTransfomer", "FlavaFusionTransfomerwithCLSToken", "MIMOTransfomer", "MIMOResNet", "MultimodalBertClf", "Model_", "FusedAdamW", "BertAdam",
           "get_cosine_schedule_with_warmup", "acc", "ops"]
