import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        from wicca_b200 import _capi
        return _capi.load().wicca_device_count() > 0
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: only skip when the user
    # did not ask for gpu tests explicitly.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def icon_golden():
    """(cases, outputs) produced by the live reference (tests/golden/make_golden.py)."""
    z = np.load(GOLDEN / "haar_icon_golden.npz")
    cases = json.loads(str(z["cases"][0]))
    return cases, [z[f"icon_{i}"] for i in range(len(cases))]


@pytest.fixture(scope="session")
def resize_golden():
    z = np.load(GOLDEN / "resize_area_golden.npz")
    cases = z["cases"].tolist()
    return cases, [z[f"out_{i}"] for i in range(len(cases))]
