import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The tests bind libu3dgpu.so (the product) and liboracle.so (the checker).  Build whichever is missing so that the suite does not
    # depend on build() having been called first; nvcc / gcc cross-compile without a GPU.  A failed build fails the tests loudly.
    import subprocess
    lib = os.path.join(ROOT, "urho3d_b200", "lib", "libu3dgpu.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "urho3d_b200", "csrc")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def ref_available():
    from oracle import refbind
    return refbind.available()


@pytest.fixture(scope="session")
def gpu_ctx():
    """One device context for the whole GPU test session.  Fails (does not skip) when the CUDA library cannot run."""
    from urho3d_b200 import capi
    ctx = capi.Context(0)
    for kv in filter(None, os.environ.get("U3D_DEBUG_OPTS", "").split(",")):  # run the suite under a tuning option, e.g. "dense_pool=1"
        name, val = kv.split("=")
        capi.debug_set_option(name, int(val))
    yield ctx
    ctx.close()


@pytest.fixture(params=["per_view", "dense"])
def cull_path(request):
    """Runs a GPU test once per cull path of the batched entry point: the one-CTA-per-view kernel (cull.cu) and the batch-wide dense
    launches (cull_dense.cu), which by default only takes batches of 16 views and more."""
    from urho3d_b200 import capi
    capi.debug_set_option("cull_dense", 0 if request.param == "per_view" else 2)
    yield request.param
    capi.debug_set_option("cull_dense", 1)
