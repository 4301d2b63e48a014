import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# tests/test_gpu_dp.py runs several emulated data-parallel ranks on ONE GPU, each with its own streams / CUDA graphs; with
# the default of 8 hardware work queues, streams of different ranks can alias onto one queue and a rank's kernels would sit
# behind another rank's spinning arrive kernel (false dependency).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: on a CPU-only machine they are SKIPPED (not failed), so a
    plain `pytest tests` shows the status of the CPU suite."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
