import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_available():
    from oracle import refbind
    return refbind.available()


@pytest.fixture(scope="session")
def gpu_ctx():
    """One device context for the whole GPU test session.  Fails (does not skip) when the CUDA library cannot run."""
    from urho3d_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
