"""Shared test helpers: small deterministic scenes and the three implementations side by side."""
import numpy as np

from urho3d_b200 import camera, scenes


def small_scene(n=3000, k=96, extent=80.0, seed=7):
    return scenes.Scene(n, k, extent, seed=seed)


def stress_views(extent, width, height, count, seed=99):
    """Cameras that provoke clipping: low, close to walls, steep pitch, plus ordinary agent cameras."""
    rs = np.random.RandomState(seed)
    out = []
    aspect = float(width) / float(height)
    for i in range(count):
        pos = [rs.uniform(-extent, extent), rs.uniform(0.3, 9.0), rs.uniform(-extent, extent)]
        out.append(camera.make_view(pos, rs.uniform(0, 360), rs.uniform(-35, 35), rs.choice([30.0, 45.0, 70.0, 100.0]), aspect,
                                    rs.choice([0.05, 0.1, 1.0]), rs.choice([60.0, 2.0 * extent, 1000.0])))
    return out


def ref_draw_view(ref, scene, view):
    """Reference, threaded-branch semantics: returns (visible ids, (submitted, candidates, visible))."""
    ref.set_view_raw(view)
    vis, counts, _ = ref.run_view(scene)
    return vis, counts


def make_ref(scene, width, height, workers=0):
    from oracle import refbind
    r = refbind.Ref(workers)
    r.octree_set_size(scene.octree_min, scene.octree_max, scene.octree_levels)
    r.add_drawables(scene.aabb, scene.flags, scene.view_mask, scene.occludee, scene.cast_shadows)
    r.ob_set_size(width, height)
    return r


def make_oracle(scene, flat, width, height):
    from oracle import oraclebind
    tree = oraclebind.OracleTree(flat, scene.aabb, scene.flags, scene.view_mask, scene.occludee, scene.cast_shadows)
    ob = oraclebind.OracleOB()
    ob.set_size(width, height)
    return tree, ob


def host_flatten(scene):
    """Flatten with the product's host octree mirror (CPU code in libu3dgpu.so)."""
    from urho3d_b200 import capi
    t = capi.Octree(scene.octree_min, scene.octree_max, scene.octree_levels)
    t.add(scene.aabb, scene.flags, scene.view_mask, scene.occludee, scene.cast_shadows)
    flat = t.flatten()
    return t, flat


def random_rays(extent, count, seed=1, stacked_boxes=None):
    """Rays for Octree::Raycast tests: from inside / outside the populated region, axis-parallel ones (zero direction components), rays
    starting inside boxes (distance 0 ties) and a few with a finite maxDistance.  Returns [(origin, direction, max_distance)]."""
    rs = np.random.RandomState(seed)
    rays = []
    for i in range(count):
        o = np.array([rs.uniform(-extent, extent), rs.uniform(0.2, 6.0), rs.uniform(-extent, extent)], np.float32)
        kind = i % 5
        if kind == 0:
            d = np.array([1.0, 0.0, 0.0], np.float32) * (1 if rs.uniform() < 0.5 else -1)
        elif kind == 1:
            d = np.array([0.0, -1.0, 0.0], np.float32)
            o[1] = 50.0
        else:
            d = rs.normal(size=3).astype(np.float32)
            d[1] *= np.float32(0.15)
            d = (d / np.float32(np.sqrt(np.float32((d * d).sum())))).astype(np.float32)
        if kind == 4 and stacked_boxes is not None:
            b = stacked_boxes[rs.randint(len(stacked_boxes))]
            o = ((b[:3] + b[3:]) * np.float32(0.5)).astype(np.float32)
        maxd = np.float32(np.inf) if i % 3 else np.float32(rs.uniform(5.0, 2.0 * extent))
        rays.append((o, d, maxd))
    return rays
