import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The C-ABI library is a build artefact (git-ignored): make sure it exists before collection.
    # (loaded by path: importing the package needs the library to exist already)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_spmf_b200_build", os.path.join(ROOT, "spmf_b200", "build.py"))
    _b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(_b)
    _b.build()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
