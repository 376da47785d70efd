import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(os.path.join(GOLDEN, name), weights_only=False)
    return load


# ---- measured-error log: bf16-path tests record what they measured (max over the session per key);
#      written to gpurun_out/measured_errors.json so that tolerances can be held to <= 3x observed
_MEASURED = {}


def record_err(key, value):
    value = float(value)
    if key not in _MEASURED or value > _MEASURED[key]:
        _MEASURED[key] = value


@pytest.fixture(scope="session")
def measured():
    return record_err


def pytest_sessionfinish(session, exitstatus):
    if _MEASURED:
        import json
        out = os.path.join(ROOT, "gpurun_out")
        try:
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "measured_errors.json"), "w") as f:
                json.dump(dict(sorted(_MEASURED.items())), f, indent=1)
        except OSError:
            pass
