import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "computer-vision-shoplifting-detection_b200"
for p in (str(REPO), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def import_dropin(which: str):
    """Import the drop-in package tree `shopformer` or `shopformer_2` the way the reference scripts
    do (top-level `models`, `data`, `utils`), isolated from the other variant."""
    import importlib
    root = PKG / which
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("models", "data", "utils")}
    sys.path.insert(0, str(root))
    try:
        mods = {"models": importlib.import_module("models"),
                "data": importlib.import_module("data.poselift_dataset"),
                "metrics": importlib.import_module("utils.metrics")}
    finally:
        sys.path.remove(str(root))
        mine = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("models", "data", "utils")}
        sys.modules.update(saved)
    mods["_modules"] = mine
    return mods


@pytest.fixture(scope="session")
def dropin1():
    return import_dropin("shopformer")


@pytest.fixture(scope="session")
def dropin2():
    return import_dropin("shopformer_2")
