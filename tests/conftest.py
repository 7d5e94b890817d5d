import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        from lineqgpr_b200 import _lib
        return _lib.lib().lqg_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Builds liboracle_tmg.so on first use and, where
    /root/reference exists, oracle/_ref/libtmg_ref.so from the unmodified reference source."""
    from oracle import lineqgpr_oracle as O
    O.build(ref=True)
    return O


@pytest.fixture(scope="session")
def golden():
    d = ROOT / "tests" / "golden"

    def load(name):
        return dict(np.load(d / f"{name}.npz"))
    return load


@pytest.fixture(scope="session")
def lqg():
    import lineqgpr_b200
    from lineqgpr_b200 import build
    build.build()
    return lineqgpr_b200


@pytest.fixture(autouse=True)
def _guard_bands_intact(request):
    """Under LQG_GUARD=1 (canary bands around every internal device buffer) each GPU test also
    asserts that none of the bands was overwritten while it ran."""
    yield
    import os
    if "gpu" in request.keywords and os.environ.get("LQG_GUARD", "0") not in ("", "0"):
        from lineqgpr_b200 import _lib
        assert _lib.lib().lqg_guard_violations() == 0
