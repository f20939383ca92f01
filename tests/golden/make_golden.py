"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref/libu3dref.so (oracle/build_ref.sh).
Run in the build container (needs /root/reference to have been compiled):  python tests/golden/make_golden.py
The fixtures hold inputs (scene seed parameters + the exact per-view matrices/planes) and the reference's outputs
(depth plane, every mip level, octree query result, visible set, numTriangles_), so they can be checked anywhere."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import helpers  # noqa: E402
from urho3d_b200 import scenes  # noqa: E402


def make(name, n, k, extent, seed, width, height, nviews, stress):
    sc = scenes.Scene(n, k, extent, seed=seed)
    ref = helpers.make_ref(sc, width, height)
    flat = ref.flatten()
    views = helpers.stress_views(extent, width, height, nviews) if stress else scenes.agent_cameras(nviews, extent, width, height)
    out = dict(n=n, k=k, extent=extent, seed=seed, width=width, height=height, stress=int(stress),
               flat_parent=flat["parent"], flat_level=flat["level"], flat_box=flat["box"], flat_first=flat["first"], flat_count=flat["count"],
               flat_order=flat["order"])
    for i, v in enumerate(views):
        vis, counts = helpers.ref_draw_view(ref, sc, v)
        out["v%d_view" % i] = v.view
        out["v%d_proj" % i] = v.proj
        out["v%d_planes" % i] = v.planes
        out["v%d_nearfar" % i] = np.array([v.near, v.far], np.float32)
        out["v%d_depth" % i] = ref.ob_depth()
        for lv, m in enumerate(ref.ob_mips()):
            out["v%d_mip%d" % (i, lv)] = m
        out["v%d_query" % i] = ref.query(1)
        out["v%d_visible" % i] = vis
        out["v%d_frustum_query" % i] = ref.query(0)
        out["v%d_shadow_query" % i] = ref.query(2)
        out["v%d_counts" % i] = counts
        out["v%d_numtris" % i] = np.array([ref.ob_num_triangles()], np.uint32)
    out["nviews"] = len(views)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    ref.close()
    print(name, "written")


if __name__ == "__main__":
    make("golden_agents_128x80", 3000, 96, 80.0, 7, 128, 80, 4, False)
    make("golden_stress_64x64", 1500, 64, 40.0, 11, 64, 64, 6, True)
    make("golden_odd_256x36", 2000, 64, 60.0, 13, 256, 35, 3, True)
