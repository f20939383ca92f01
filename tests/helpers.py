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


def shadow_scenarios(ref, extent, count, seed=5, shared_main=False):
    """Main (cull) camera + shadow camera pairs for View::ProcessShadowCasters tests: directional cascades (orthographic shadow camera, light type
    0), spot lights (perspective, type 1) and point-light cube faces (90 degree perspective, type 2).  The split inputs come from the reference's
    own Camera / Frustum code (oracle/_ref: ref.shadow_split_setup), so they are plain inputs for every implementation.
    Returns a list of dicts: split (see refbind.shadow_split_setup), light_type, main_view (12), main_pos (3), main view object."""
    from urho3d_b200 import camera
    rs = np.random.RandomState(seed)
    out = []
    main = None
    for i in range(count):
        if main is None or not shared_main:  # one frame has one cull camera: shared_main keeps it for every split
            main = ([rs.uniform(-extent, extent) * 0.5, rs.uniform(1.0, 6.0), rs.uniform(-extent, extent) * 0.5], rs.uniform(0, 360),
                    rs.uniform(-5, 25), float(rs.choice([120.0, 300.0])))
        mpos, myaw, mpitch, mfar = main
        mq = camera.quat_from_euler(mpitch, myaw, 0.0)
        main_cam = np.array(list(mpos) + list(mq) + [45.0, 0.1, mfar, 0.0, 20.0], np.float32)
        mv = camera.make_view(mpos, myaw, mpitch, 45.0, 1.0, 0.1, mfar)
        kind = i % 3
        if kind == 0:    # directional cascade: orthographic camera above the scene looking down at an angle
            spos = [mpos[0] + rs.uniform(-30, 30), rs.uniform(80.0, 200.0), mpos[2] + rs.uniform(-30, 30)]
            sq = camera.quat_from_euler(rs.uniform(40, 80), rs.uniform(0, 360), 0.0)
            shadow_cam = np.array(list(spos) + list(sq) + [45.0, 0.0, 400.0, 1.0, float(rs.choice([40.0, 120.0, 300.0]))], np.float32)
            near_z, far_z = float(rs.choice([0.0, 10.0])), float(rs.choice([40.0, 150.0, 1000.0]))
        elif kind == 1:  # spot light
            spos = [mpos[0] + rs.uniform(-40, 40), rs.uniform(3.0, 25.0), mpos[2] + rs.uniform(-40, 40)]
            sq = camera.quat_from_euler(rs.uniform(10, 70), rs.uniform(0, 360), 0.0)
            shadow_cam = np.array(list(spos) + list(sq) + [float(rs.choice([30.0, 60.0])), 0.5, float(rs.uniform(30, 90)), 0.0, 20.0], np.float32)
            near_z, far_z = 0.0, float(rs.choice([60.0, 1000.0]))
        else:            # one face of a point light's cube
            spos = [mpos[0] + rs.uniform(-40, 40), rs.uniform(1.0, 8.0), mpos[2] + rs.uniform(-40, 40)]
            pitch, yaw = [(0.0, 90.0), (0.0, -90.0), (-90.0, 0.0), (90.0, 0.0), (0.0, 0.0), (0.0, 180.0)][rs.randint(6)]
            sq = camera.quat_from_euler(pitch, yaw, 0.0)
            rng = float(rs.uniform(20, 60))
            shadow_cam = np.array(list(spos) + list(sq) + [90.0, 0.01 * rng, rng, 0.0, 20.0], np.float32)
            near_z, far_z = 0.0, 1000.0
        split = ref.shadow_split_setup(main_cam, shadow_cam, near_z, far_z)
        if split["degenerate"]:
            continue
        out.append(dict(split=split, light_type=kind, main_view=mv.view.reshape(-1).copy(), main_pos=np.array(mpos, np.float32), view=mv,
                        main_cam=main_cam, shadow_cam=shadow_cam, near_z=near_z, far_z=far_z))
    return out
