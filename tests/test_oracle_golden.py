"""The C oracle (oracle/u3d_oracle.c) against golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import oraclebind
from urho3d_b200 import camera, scenes

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_views(g):
    views = []
    for i in range(int(g["nviews"])):
        nf = g["v%d_nearfar" % i]
        views.append(camera.CameraView(g["v%d_view" % i], g["v%d_proj" % i], g["v%d_planes" % i], nf[0], nf[1]))
    return views


def load_scene(g):
    return scenes.Scene(int(g["n"]), int(g["k"]), float(g["extent"]), seed=int(g["seed"]))


def golden_flat(g):
    return dict(parent=g["flat_parent"], level=g["flat_level"], box=g["flat_box"], first=g["flat_first"], count=g["flat_count"], order=g["flat_order"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    assert GOLDEN, "golden fixtures missing"
    g = np.load(path)
    sc = load_scene(g)
    w, h = int(g["width"]), int(g["height"])
    tree, ob = helpers.make_oracle(sc, golden_flat(g), w, h)
    for i, v in enumerate(load_views(g)):
        sub = ob.draw_scene_view(sc, v)
        assert sub == int(g["v%d_counts" % i][0])
        assert np.array_equal(ob.depth(), g["v%d_depth" % i]), "depth plane differs from the reference"
        for lv, m in enumerate(ob.mips()):
            assert np.array_equal(m, g["v%d_mip%d" % (i, lv)]), "mip level %d differs" % lv
        assert ob.num_triangles() == int(g["v%d_numtris" % i][0])
        cand = tree.query(oraclebind.lib() and 1, v.planes, ob, as_ids=False)
        assert np.array_equal(tree.order[cand], g["v%d_query" % i])
        vis = tree.order[tree.check_visibility(ob, cand)]
        assert np.array_equal(vis, g["v%d_visible" % i])
        assert np.array_equal(tree.query(0, v.planes, None), g["v%d_frustum_query" % i])
        assert np.array_equal(tree.query(2, v.planes, None), g["v%d_shadow_query" % i])
    ob.close()


def test_golden_has_clipping_and_occlusion():
    """The fixtures must exercise what they claim: culled candidates, clipped triangles, odd height rounding."""
    g = np.load([p for p in GOLDEN if "stress" in p][0])
    culled = sum(len(g["v%d_query" % i]) - len(g["v%d_visible" % i]) for i in range(int(g["nviews"])))
    assert culled > 0
    g2 = np.load([p for p in GOLDEN if "odd" in p][0])
    assert g2["v0_depth"].shape == (36, 256)  # SetSize rounds 35 up (OcclusionBuffer.cpp:63-65)
