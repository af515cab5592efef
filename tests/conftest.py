import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


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
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden.npz")))


@pytest.fixture(scope="session")
def pipeline_cfg():
    from slide_b200 import weights
    return weights.load_json("pipeline_airplane.json")


@pytest.fixture(autouse=True)
def _default_kernel_tuning(request):
    """libslide_b200.so reads its SLIDE_TC_* / SLIDE_PAIR_* knobs once; GPU tests that override them (monkeypatch) call
    slide_tc_reload_tuning() themselves -- this restores the default dispatch for whatever test runs next."""
    yield
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            from slide_b200 import lib
            os.environ.pop("SLIDE_TC_PERSIST_MIN_TILES", None)
            os.environ.pop("SLIDE_TC_PERSIST_MIN_K", None)
            os.environ.pop("SLIDE_TC_PERSIST_GRID", None)
            lib.load().slide_tc_reload_tuning()
