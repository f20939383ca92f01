"""The product's host octree mirror (CPU code inside libu3dgpu.so) must keep exactly the reference Octree's structure."""
import numpy as np
import pytest

import helpers
from oracle import refbind
from urho3d_b200 import capi, scenes

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libu3dref.so not built")


def assert_same_flat(a, b):
    for key in ("parent", "level", "first", "count", "order"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["box"], b["box"]), "culling boxes differ bitwise"


@pytest.mark.parametrize("n,extent,levels,size", [(5000, 150.0, 8, 1000.0), (3000, 900.0, 8, 1000.0), (2000, 1500.0, 5, 1000.0), (1000, 50.0, 3, 64.0)])
def test_flatten_matches_reference(n, extent, levels, size):
    sc = scenes.Scene(n, 4, extent, seed=n, octree_size=size, octree_levels=levels)
    # a few huge and a few out-of-bounds drawables: they must stay in the root (Octree.cpp:141-142)
    sc.aabb[:5, 3:] += 700.0
    sc.aabb[5:8] += 5000.0
    ref = helpers.make_ref(sc, 64, 64)
    t, flat = helpers.host_flatten(sc)
    assert_same_flat(flat, ref.flatten())
    assert flat["count"].sum() == n
    t.close()
    ref.close()


def test_default_octree_shape_of_the_million_scene_sample():
    """drawables of the C2/C3 generator land in the deepest level; non-occludees (1 %) stay in the root."""
    sc = scenes.Scene(20000, 4, 500.0)
    t, flat = helpers.host_flatten(sc)
    assert flat["count"][0] == int((sc.occludee == 0).sum())
    assert flat["level"].max() == 8
    t.close()
