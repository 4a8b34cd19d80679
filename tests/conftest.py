import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native_lib():
    """The built C-ABI library; builds it on first use (nvcc cross-compiles without a GPU)."""
    from tuatara_b200 import _native, build

    if not _native.LIB_PATH.exists():
        build.build()
    return _native.lib()


@pytest.fixture(scope="session")
def oracle_models():
    """Seeded random-init fp32 oracle networks (checkpoints are unavailable offline)."""
    from oracle.models import make_craft, make_parseq

    return make_craft(0), make_parseq("base", 0)


@pytest.fixture(scope="session")
def weights_dir(oracle_models):
    """craft.ttw / parseq.ttw exported from the oracle models (cached under tests/_cache)."""
    from tuatara_b200 import weights

    d = ROOT / "tests" / "_cache" / "weights_seed0"
    d.mkdir(parents=True, exist_ok=True)
    craft, parseq = oracle_models
    if not (d / "craft.ttw").exists():
        weights.export_craft(craft.state_dict(), d / "craft.ttw")
    if not (d / "parseq.ttw").exists():
        weights.export_parseq(parseq.state_dict(), d / "parseq.ttw")
    return str(d)


@pytest.fixture(scope="session")
def engine(native_lib, weights_dir):
    import tuatara_b200 as tb

    e = tb.Engine(weights_dir)
    yield e
    e.close()
