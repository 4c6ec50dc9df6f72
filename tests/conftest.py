import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) when no device is visible, e.g. a bare `pytest tests/` here.
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="module", params=["gather", "scatter", "cluster", "auto"])
def pileup_kernel(request):
    """Every engine test runs against both hot-kernel formulations (k_pileup5 gather, k_pileup7 scatter, k_pileup7c cluster scatter) and the
    engine's own per-region choice.  PB_PILEUP is read by pb_create, so engines made inside the module see it."""
    old = os.environ.get("PB_PILEUP")
    val = {"gather": "5", "scatter": "7", "cluster": "8", "auto": None}[request.param]
    if val is None:
        os.environ.pop("PB_PILEUP", None)
    else:
        os.environ["PB_PILEUP"] = val
    yield request.param
    if old is None:
        os.environ.pop("PB_PILEUP", None)
    else:
        os.environ["PB_PILEUP"] = old
