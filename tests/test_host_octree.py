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


def test_moves_and_removals_match_reference():
    """Octree::Update's reinsertion rule (Octree.cpp:440-472), AddDrawable-before-RemoveDrawable ordering and empty-branch deletion
    (Octree.h:123-136): after random moves, resizes and removals the mirror must still flatten to the reference's arrays."""
    sc = scenes.Scene(3000, 4, 200.0, seed=77)
    ref = helpers.make_ref(sc, 64, 64)
    t, flat = helpers.host_flatten(sc)
    assert_same_flat(flat, ref.flatten())
    # every new drawable queues itself once (Drawable::OnMarkedDirty at creation); an engine drains that queue in its first frame
    ref.move_drawables(np.zeros(0, np.int32), np.zeros((0, 6), np.float32))
    assert_same_flat(flat, ref.flatten())
    rs = np.random.RandomState(8)
    boxes = sc.aabb.copy()
    removed = set()
    for step in range(6):
        ids = [i for i in rs.choice(sc.n, size=400, replace=False) if i not in removed]
        kind = rs.randint(0, 4, size=len(ids))
        for i, k in zip(ids, kind):
            c = 0.5 * (boxes[i, :3] + boxes[i, 3:])
            e = 0.5 * (boxes[i, 3:] - boxes[i, :3])
            if k == 0:
                c = c + rs.uniform(-0.5, 0.5, 3)        # small move: usually stays in its octant
            elif k == 1:
                c = c + rs.uniform(-60, 60, 3) * [1, 0.05, 1]  # big move
            elif k == 2:
                e = e * rs.uniform(0.2, 12.0)            # resize: changes the level it fits
            else:
                c = c + rs.uniform(-3000, 3000, 3)       # may leave the octree bounds -> root
            boxes[i, :3], boxes[i, 3:] = (c - e).astype(np.float32), (c + e).astype(np.float32)
        ref.move_drawables(ids, boxes[ids])
        for i in ids:
            t.update(int(i), boxes[i])
        for i in rs.choice(sc.n, size=40, replace=False):
            if i not in removed:
                removed.add(int(i))
                ref.remove_drawable(int(i))
                t.remove(int(i))
        a, b = t.flatten(), ref.flatten()
        # the reference harness reports a flat `order` padded to the original count; compare what the tree holds
        nb = int(b["count"].sum())
        b = dict(b, order=b["order"][:nb])
        assert_same_flat(a, b)
        assert nb == sc.n - len(removed)
    t.close()
    ref.close()


def test_host_sort_follows_the_reference_sort():
    """u3d_sort_by_key (host code of the product: orders ray hits and, later, occluders) leaves ties exactly where the reference's Sort does."""
    ref = refbind.Ref(0)
    rs = np.random.RandomState(11)
    for n in list(range(0, 36)) + [64, 333, 4096]:
        for distinct in (2, 9, 10 ** 6):
            keys = rs.randint(0, distinct, size=n).astype(np.float32)
            keys[rs.uniform(size=n) < 0.05] = np.inf
            vals = np.arange(n, dtype=np.int32)
            rk, rv = ref.sort_by_key(keys, vals)
            gk, gv = capi.sort_by_key(keys, vals.astype(np.uint32))
            assert np.array_equal(rk, gk) and np.array_equal(rv.astype(np.uint32), gv), (n, distinct)
    ref.close()


def test_wild_drawable_boxes_keep_the_reference_structure():
    """Drawables whose boxes are inverted, flat, points, enormous, infinite or NaN: the mirror's insertion rule (CheckDrawableFit, Octree.cpp:
    134-190) must put every one of them where the reference's Octree does, and the flattened arrays must stay bit-equal."""
    rs = np.random.RandomState(8)
    n = 3000
    sc = scenes.Scene(n, 4, 200.0, seed=31)
    b = sc.aabb
    k = n // 12
    b[0 * k:1 * k] = b[0 * k:1 * k][:, [3, 4, 5, 0, 1, 2]]        # inverted (undefined BoundingBox)
    b[1 * k:2 * k, 3:] = b[1 * k:2 * k, :3]                        # points
    b[2 * k:3 * k, 4] = b[2 * k:3 * k, 1]                          # flat
    b[3 * k:4 * k, :3] -= rs.uniform(0, 3000, size=(k, 3)).astype(np.float32)   # from a little to far larger than the octree
    b[3 * k:4 * k, 3:] += rs.uniform(0, 3000, size=(k, 3)).astype(np.float32)
    b[4 * k:5 * k] += np.float32(990.0)                            # straddling / outside the octree bounds (+-1000)
    b[5 * k:6 * k, 3] = np.inf
    b[6 * k:7 * k, 2] = -np.inf
    b[7 * k:8 * k, 0] = np.nan
    b[8 * k:9 * k, 5] = np.nan
    b[9 * k:10 * k] = np.float32(1e30)
    ref = helpers.make_ref(sc, 64, 64)
    t, flat = helpers.host_flatten(sc)
    assert_same_flat(flat, ref.flatten())
    t.close()
    ref.close()
