import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the oracle's C restatement is test infrastructure: build it on demand (gcc, <1 s)
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle_solver.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
