import os
import sys

import pytest

# parity tests build randomly initialised models and load oracle.weights state_dicts into them: the explicit opt-in that
# tris_b200.clip_model.load requires when no CLIP checkpoint file is present (see its docstring)
os.environ.setdefault("TRIS_ALLOW_RANDOM_INIT", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "stage1_golden.npz"), allow_pickle=False)
