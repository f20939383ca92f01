"""libu3dgpu.so loads without a GPU and exports every symbol include/u3d_gpu.h declares; host-only entry points behave."""
import ctypes
import os
import re

import numpy as np
import pytest

from urho3d_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "u3d_gpu.h")).read()
    return sorted(set(re.findall(r"U3D_API\s+[\w\s\*]+?\b(u3d_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 60
    L = ctypes.CDLL(capi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libu3dgpu.so does not export " + n
    assert sorted(capi.SIGNATURES) == names, "capi.SIGNATURES and include/u3d_gpu.h disagree"


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.U3DError) as e:
        capi.Context(0)
    assert e.value.code == 5 and "no CPU path" in str(e.value)


def test_octree_argument_errors():
    with pytest.raises(capi.U3DError):
        t = capi.Octree()
        t.update(5, np.zeros(6, np.float32))
