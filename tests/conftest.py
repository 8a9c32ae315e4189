import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    # A/B builds (tools/build_variant.py): run the SAME tests against a variant library. Test infrastructure only — the
    # package itself always loads bagel_b200/libbagel_b200.so.
    lib = os.environ.get("BAGEL_TEST_LIB")
    if lib:
        from pathlib import Path
        from bagel_b200 import _cabi
        _cabi.LIB_PATH = Path(lib).resolve()


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
