import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _emulated():
    """SWIM_TEST_EMU=1: run the `-m gpu` tests in this container against tests/emu/libswim_emu.so (the CUDA sources on
    the SIMT emulator). A developer's dry run of the GPU suite — sizes that only make sense on hardware
    are slow, multi-process tests still skip — never what the GPU box does."""
    return os.environ.get("SWIM_TEST_EMU") == "1"


def pytest_sessionstart(session):
    if _emulated():
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        import swim_b200._lib as L
        L.SO_PATH, L._lib = build_emu.build(), None


def pytest_collection_modifyitems(config, items):
    if _has_gpu() or _emulated():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
