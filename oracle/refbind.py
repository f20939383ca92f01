"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libu3dref.so = the UNMODIFIED reference classes
(OcclusionBuffer, Octree, Camera, FrustumOctreeQuery ...) compiled from /root/reference by oracle/build_ref.sh,
driven through oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
--impl reference legs may import this; the product (urho3d_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libu3dref.so")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_ubyte)
_u32 = C.POINTER(C.c_uint)


def available():
    return os.path.exists(LIB_PATH)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class Ref:
    """One reference Scene+Octree+Camera+OcclusionBuffer."""

    def __init__(self, worker_threads=0):
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int]
        for name in ("ref_destroy", "ref_ob_set_view", "ref_ob_clear", "ref_ob_reset", "ref_ob_draw", "ref_ob_build_hierarchy"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.ref_octree_set_size.argtypes = [C.c_void_p, _f, _f, C.c_uint]
        L.ref_add_drawables.argtypes = [C.c_void_p, C.c_int, _f, _u8, _u32, _u8, _u8]
        L.ref_octree_num_octants.argtypes = [C.c_void_p]
        L.ref_octree_flatten.argtypes = [C.c_void_p, _i, _i, _f, _i, _i, _i]
        L.ref_camera_set.argtypes = [C.c_void_p, _f, _f, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, C.c_int]
        L.ref_camera_get.argtypes = [C.c_void_p, _f, _f, _f, _f, _f, _i]
        L.ref_set_frustum_raw.argtypes = [C.c_void_p, _f]
        L.ref_ob_set_view_raw.argtypes = [C.c_void_p, _f, _f, C.c_float, C.c_float, C.c_int]
        L.ref_ob_get_viewproj.argtypes = [C.c_void_p, _f]
        L.ref_frustum_is_inside.argtypes = [C.c_void_p, _f, C.c_int]
        L.ref_ob_set_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_ob_width.argtypes = [C.c_void_p]
        L.ref_ob_height.argtypes = [C.c_void_p]
        L.ref_ob_set_max_triangles.argtypes = [C.c_void_p, C.c_uint]
        L.ref_ob_set_cull_mode.argtypes = [C.c_void_p, C.c_int]
        L.ref_ob_get_cull_mode.argtypes = [C.c_void_p]
        L.ref_ob_add_triangles.argtypes = [C.c_void_p, _f, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
        L.ref_ob_num_triangles.argtypes = [C.c_void_p]
        L.ref_ob_num_triangles.restype = C.c_uint
        L.ref_ob_is_visible.argtypes = [C.c_void_p, _f]
        L.ref_ob_is_visible_many.argtypes = [C.c_void_p, C.c_int, _f, _u8]
        L.ref_ob_read_depth.argtypes = [C.c_void_p, _i]
        L.ref_ob_num_mips.argtypes = [C.c_void_p]
        L.ref_ob_read_mip.argtypes = [C.c_void_p, C.c_int, _i]
        L.ref_query.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_uint, _i, C.c_int]
        L.ref_check_visibility.argtypes = [C.c_void_p, C.c_int, _i, C.c_int]
        L.ref_run_view.argtypes = [C.c_void_p, C.c_int, _f, _f, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_int,
                                   C.c_uint, C.c_uint, _i, C.c_int, _i, C.POINTER(C.c_double)]
        L.ref_num_threads.argtypes = [C.c_void_p]
        L.ref_query_shape.argtypes = [C.c_void_p, C.c_int, _f, C.c_uint, C.c_uint, _i, C.c_int]
        L.ref_move_drawables.argtypes = [C.c_void_p, C.c_int, _i, _f]
        L.ref_remove_drawable.argtypes = [C.c_void_p, C.c_int]
        L.ref_select_occluders.argtypes = [C.c_void_p, C.c_int, _f, _u32, _f, _f, _f, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _i, _f, C.c_int]
        L.ref_select_occluders_text.argtypes = [C.c_void_p, C.c_int, _f, _u32, _f, C.c_float, _i, _f, C.c_int]
        L.ref_run_view_selected.argtypes = [C.c_void_p, C.c_int, _i, _f, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_uint,
                                            C.c_uint, C.c_uint, _i, C.c_int, _i]
        L.ref_query_zone_occluder.argtypes = [C.c_void_p, _u8, C.c_uint, _i, C.c_int]
        L.ref_shadow_split_setup.argtypes = [C.c_void_p, _f, _f, C.c_float, C.c_float, _f, _f, _f, _f, _f]
        L.ref_process_shadow_casters.argtypes = [C.c_void_p, C.c_int, _i, _f, _f, _f, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint, _u32, _f, _f, _u8,
                                                 _f, _f, C.c_int, _i]
        L.ref_process_shadow_casters_text.argtypes = [C.c_void_p, C.c_int, _i, _f, _f, C.c_float, C.c_float, C.c_int, C.c_uint, _u32, _f, _f, _u8, _i]
        L.ref_run_view_serial.argtypes = [C.c_void_p, C.c_int, _i, _f, _f, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_int, C.c_uint,
                                          C.c_uint, C.c_uint, _i, C.c_int, _i]
        L.ref_raycast.argtypes = [C.c_void_p, _f, C.c_uint, C.c_uint, C.c_int, _i, _f, _f, C.c_int]
        L.ref_sort_by_key.argtypes = [C.c_int, _f, _i]
        L.ref_check_in_view.argtypes = [C.c_void_p, C.c_int, _f, _f, C.c_int, _f, _i, C.c_int, _f]
        L.ref_check_in_view_text.argtypes = [C.c_void_p, C.c_int, _f, _i, C.c_int, _f]
        self.h = L.ref_create(int(worker_threads))
        self._keep = []
        self.n = 0

    def close(self):
        if self.h:
            self.lib.ref_destroy(self.h)
            self.h = None

    # --- octree -------------------------------------------------------------------------------
    def octree_set_size(self, mn, mx, levels):
        mn = np.ascontiguousarray(mn, dtype=np.float32)
        mx = np.ascontiguousarray(mx, dtype=np.float32)
        self.lib.ref_octree_set_size(self.h, _p(mn, _f), _p(mx, _f), int(levels))

    def add_drawables(self, aabb6, flags=None, view_mask=None, occludee=None, cast_shadows=None):
        aabb6 = np.ascontiguousarray(aabb6, dtype=np.float32)
        n = aabb6.shape[0]
        arrs = [None if a is None else np.ascontiguousarray(a) for a in (flags, view_mask, occludee, cast_shadows)]
        self.lib.ref_add_drawables(self.h, n, _p(aabb6, _f), _p(arrs[0], _u8), _p(arrs[1], _u32), _p(arrs[2], _u8), _p(arrs[3], _u8))
        self.n += n

    def move_drawables(self, ids, aabb6):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        a = np.ascontiguousarray(aabb6, dtype=np.float32)
        self.lib.ref_move_drawables(self.h, len(ids), _p(ids, _i), _p(a, _f))

    def remove_drawable(self, idx):
        self.lib.ref_remove_drawable(self.h, int(idx))

    def flatten(self):
        """DFS pre-order dump: parent, level, cullingBox(6), first, count per octant; order = drawable ids in DFS order."""
        m = self.lib.ref_octree_num_octants(self.h)
        parent = np.zeros(m, np.int32)
        level = np.zeros(m, np.int32)
        box = np.zeros((m, 6), np.float32)
        first = np.zeros(m, np.int32)
        count = np.zeros(m, np.int32)
        order = np.zeros(max(self.n, 1), np.int32)
        self.lib.ref_octree_flatten(self.h, _p(parent, _i), _p(level, _i), _p(box, _f), _p(first, _i), _p(count, _i), _p(order, _i))
        return dict(parent=parent, level=level, box=box, first=first, count=count, order=order[:self.n])

    # --- camera / view ------------------------------------------------------------------------
    def camera_set(self, pos, quat_wxyz, fov, aspect, near, far, ortho=False, ortho_size=20.0, zoom=1.0, flip=False):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        q = np.ascontiguousarray(quat_wxyz, dtype=np.float32)
        self.lib.ref_camera_set(self.h, _p(pos, _f), _p(q, _f), fov, aspect, near, far, int(ortho), ortho_size, zoom, int(flip))

    def camera_get(self):
        view = np.zeros((3, 4), np.float32)
        proj = np.zeros((4, 4), np.float32)
        planes = np.zeros((6, 4), np.float32)
        absn = np.zeros((6, 3), np.float32)
        nf = np.zeros(2, np.float32)
        rev = C.c_int(0)
        self.lib.ref_camera_get(self.h, _p(view, _f), _p(proj, _f), _p(planes, _f), _p(absn, _f), _p(nf, _f), C.byref(rev))
        return view, proj, planes, absn, nf, rev.value

    def set_view_raw(self, v):
        """Feed a urho3d_b200.camera.CameraView byte-for-byte: OcclusionBuffer::SetView fields + query frustum."""
        self.lib.ref_ob_set_view_raw(self.h, _p(v.view, _f), _p(v.proj, _f), v.near, v.far, int(v.reverse_culling))
        self.lib.ref_set_frustum_raw(self.h, _p(v.planes, _f))

    def viewproj(self):
        out = np.zeros((4, 4), np.float32)
        self.lib.ref_ob_get_viewproj(self.h, _p(out, _f))
        return out

    def frustum_is_inside(self, aabb6, fast=False):
        b = np.ascontiguousarray(aabb6, dtype=np.float32)
        return self.lib.ref_frustum_is_inside(self.h, _p(b, _f), int(fast))

    # --- occlusion buffer ---------------------------------------------------------------------
    def ob_set_size(self, w, h, threaded=False):
        return bool(self.lib.ref_ob_set_size(self.h, w, h, int(threaded)))

    def ob_size(self):
        return self.lib.ref_ob_width(self.h), self.lib.ref_ob_height(self.h)

    def ob_set_view_camera(self):
        self.lib.ref_ob_set_view(self.h)

    def ob_set_max_triangles(self, n):
        self.lib.ref_ob_set_max_triangles(self.h, n)

    def ob_set_cull_mode(self, mode):
        self.lib.ref_ob_set_cull_mode(self.h, mode)

    def ob_get_cull_mode(self):
        return self.lib.ref_ob_get_cull_mode(self.h)

    def ob_clear(self):
        self.lib.ref_ob_clear(self.h)
        self._keep = []

    def ob_reset(self):
        self.lib.ref_ob_reset(self.h)

    def ob_add_triangles(self, model12, vtx, vtx_size, idx=None, start=0, count=0):
        model12 = np.ascontiguousarray(model12, dtype=np.float32)
        self._keep += [model12, vtx, idx]  # the reference keeps raw pointers until DrawTriangles
        idx_size = 0 if idx is None else idx.dtype.itemsize
        return bool(self.lib.ref_ob_add_triangles(self.h, _p(model12, _f), vtx.ctypes.data, vtx_size,
                                                  None if idx is None else idx.ctypes.data, idx_size, start, count))

    def ob_draw(self):
        self.lib.ref_ob_draw(self.h)
        self._keep = []

    def ob_build_hierarchy(self):
        self.lib.ref_ob_build_hierarchy(self.h)

    def ob_num_triangles(self):
        return self.lib.ref_ob_num_triangles(self.h)

    def ob_is_visible(self, aabb6):
        b = np.ascontiguousarray(aabb6, dtype=np.float32)
        return bool(self.lib.ref_ob_is_visible(self.h, _p(b, _f)))

    def ob_is_visible_many(self, aabb6):
        b = np.ascontiguousarray(aabb6, dtype=np.float32)
        out = np.zeros(b.shape[0], np.uint8)
        self.lib.ref_ob_is_visible_many(self.h, b.shape[0], _p(b, _f), _p(out, _u8))
        return out

    def ob_depth(self):
        w, h = self.ob_size()
        out = np.zeros((h, w), np.int32)
        self.lib.ref_ob_read_depth(self.h, _p(out, _i))
        return out

    def ob_mips(self):
        w, h = self.ob_size()
        res = []
        for lv in range(self.lib.ref_ob_num_mips(self.h)):
            w = (w + 1) // 2
            h = (h + 1) // 2
            out = np.zeros((h, w, 2), np.int32)
            self.lib.ref_ob_read_mip(self.h, lv, _p(out, _i))
            res.append(out)
        return res

    # --- queries ------------------------------------------------------------------------------
    def query(self, mode, flags=0xFF, view_mask=0xFFFFFFFF):
        """mode 0 FrustumOctreeQuery, 1 OccludedFrustumOctreeQuery, 2 ShadowCasterOctreeQuery -> drawable ids (DFS order)."""
        out = np.zeros(max(self.n, 1), np.int32)
        n = self.lib.ref_query(self.h, mode, flags, view_mask, _p(out, _i), out.shape[0])
        return out[:n].copy()

    def query_shape(self, shape, params, flags=0xFF, view_mask=0xFFFFFFFF):
        """shape 3 SphereOctreeQuery (centre.xyz, radius), 4 BoxOctreeQuery (min, max), 5 PointOctreeQuery (xyz)."""
        p = np.zeros(8, np.float32)
        fp = np.asarray(params, np.float32).reshape(-1)
        p[:fp.shape[0]] = fp
        out = np.zeros(max(self.n, 1), np.int32)
        n = self.lib.ref_query_shape(self.h, shape, _p(p, _f), flags, view_mask, _p(out, _i), out.shape[0])
        return out[:n].copy()

    def select_occluders(self, scene, view, size_threshold, draw_dist=None, num_triangles=None):
        """View::UpdateOccluders over the scene's occluder table for the frustum last set.  Returns (occluder indices, sort values), sorted."""
        k = scene.k
        tris = np.full(k, scene.mesh_idx.shape[0] // 3, np.uint32) if num_triangles is None else np.ascontiguousarray(num_triangles, np.uint32)
        dd = None if draw_dist is None else np.ascontiguousarray(draw_dist, np.float32)
        aabb = np.ascontiguousarray(scene.occ_aabb, np.float32)
        v12 = np.ascontiguousarray(view.view, np.float32)
        cp = np.ascontiguousarray(view.position, np.float32)
        out = np.zeros(max(k, 1), np.int32)
        keys = np.zeros(max(k, 1), np.float32)
        n = self.lib.ref_select_occluders(self.h, k, _p(aabb, _f), _p(tris, _u32), _p(dd, _f), _p(v12, _f), _p(cp, _f), int(view.ortho),
                                          view.half_view_size, view.ortho_size, view.far, size_threshold, _p(out, _i), _p(keys, _f), out.shape[0])
        return out[:n].copy(), keys[:n].copy()

    def select_occluders_text(self, scene, size_threshold, draw_dist=None, num_triangles=None):
        """View::UpdateOccluders compiled from the reference's own text (View.cpp:2171-2229) over real Drawables, with the camera last given to
        camera_set() as the cull camera and the frustum last set as the occluder query.  Returns (occluder indices, sort values), sorted."""
        k = scene.k
        tris = np.full(k, scene.mesh_idx.shape[0] // 3, np.uint32) if num_triangles is None else np.ascontiguousarray(num_triangles, np.uint32)
        dd = None if draw_dist is None else np.ascontiguousarray(draw_dist, np.float32)
        aabb = np.ascontiguousarray(scene.occ_aabb, np.float32)
        out = np.zeros(max(k, 1), np.int32)
        keys = np.zeros(max(k, 1), np.float32)
        n = self.lib.ref_select_occluders_text(self.h, k, _p(aabb, _f), _p(tris, _u32), _p(dd, _f), size_threshold, _p(out, _i), _p(keys, _f), out.shape[0])
        return out[:n].copy(), keys[:n].copy()

    def run_view_selected(self, scene, selected, max_triangles, flags=0xFF, view_mask=0xFFFFFFFF):
        """DrawOccluders' threaded branch over a sorted occluder list with the triangle budget, then query + predicate.
        Returns (visible ids, (submitted tris, query size, visible, activeOccluders))."""
        out = np.zeros(max(self.n, 1), np.int32)
        counts = np.zeros(4, np.int32)
        sel = np.ascontiguousarray(selected, np.int32)
        model = np.ascontiguousarray(scene.occ_model, np.float32)
        vtx = np.ascontiguousarray(scene.mesh_vtx)
        idx = np.ascontiguousarray(scene.mesh_idx)
        n = self.lib.ref_run_view_selected(self.h, len(sel), _p(sel, _i), _p(model, _f), vtx.ctypes.data, scene.mesh_stride, idx.ctypes.data,
                                           idx.itemsize, idx.shape[0], scene.cull_mode, int(max_triangles), flags, view_mask, _p(out, _i), out.shape[0],
                                           _p(counts, _i))
        return out[:n].copy(), counts

    def query_zone_occluder(self, is_occluder, view_mask=0xFFFFFFFF):
        """ZoneOccluderOctreeQuery over the frustum last set; is_occluder per drawable id."""
        occ = np.ascontiguousarray(is_occluder, np.uint8)
        out = np.zeros(max(self.n, 1), np.int32)
        n = self.lib.ref_query_zone_occluder(self.h, _p(occ, _u8), view_mask, _p(out, _i), out.shape[0])
        return out[:n].copy()

    def shadow_split_setup(self, main_cam, shadow_cam, near_z, far_z):
        """main_cam / shadow_cam: 12 floats (pos3, rot4 wxyz, fov, near, far, ortho, orthoSize), aspect 1.  Returns a dict with the split inputs of
        View::ProcessShadowCasters computed by the reference's Camera / Frustum code."""
        m = np.ascontiguousarray(main_cam, np.float32)
        sh = np.ascontiguousarray(shadow_cam, np.float32)
        lv = np.zeros(12, np.float32)
        sp = np.zeros(24, np.float32)
        lp = np.zeros(24, np.float32)
        mz = np.zeros(1, np.float32)
        fc = np.zeros(1, np.float32)
        deg = self.lib.ref_shadow_split_setup(self.h, _p(m, _f), _p(sh, _f), near_z, far_z, _p(lv, _f), _p(sp, _f), _p(lp, _f), _p(mz, _f), _p(fc, _f))
        return dict(light_view=lv, shadow_planes=sp, lvf_planes=lp, lvf_max_z=float(mz[0]), shadow_far=float(fc[0]), degenerate=bool(deg),
                    shadow_ortho=bool(sh[10] != 0))

    def process_shadow_casters(self, ids, split, light_type, light_mask, shadow_mask, shadow_distance, draw_distance, in_view, main_view, main_pos,
                               main_ortho=False):
        """The per-drawable loop of View::ProcessShadowCasters (View.cpp:2409-2450) over `ids`; returns the shadow casters in order."""
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(max(len(ids), 1), np.int32)
        sm = np.ascontiguousarray(shadow_mask, np.uint32)
        sd = np.ascontiguousarray(shadow_distance, np.float32)
        dd = np.ascontiguousarray(draw_distance, np.float32)
        iv = np.ascontiguousarray(in_view, np.uint8)
        mv = np.ascontiguousarray(main_view, np.float32).reshape(-1)
        mp = np.ascontiguousarray(main_pos, np.float32)
        n = self.lib.ref_process_shadow_casters(self.h, len(ids), _p(ids, _i), _p(split["light_view"], _f), _p(split["shadow_planes"], _f),
                                                _p(split["lvf_planes"], _f), split["lvf_max_z"], split["shadow_far"], int(split["shadow_ortho"]),
                                                light_type, light_mask, _p(sm, _u32), _p(sd, _f), _p(dd, _f), _p(iv, _u8), _p(mv, _f), _p(mp, _f),
                                                int(main_ortho), _p(out, _i))
        return out[:n].copy()

    def process_shadow_casters_text(self, ids, main_cam, shadow_cam, near_z, far_z, light_type, light_mask, shadow_mask, shadow_distance,
                                    draw_distance, in_view):
        """View::ProcessShadowCasters compiled from the reference's own text (View.cpp:2379-2453 + IsShadowCasterVisible) for one split, with real
        cull / shadow Cameras built from the 12-float descriptions of shadow_split_setup.  Returns the shadow casters in order."""
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(max(len(ids), 1), np.int32)
        m = np.ascontiguousarray(main_cam, np.float32)
        sh = np.ascontiguousarray(shadow_cam, np.float32)
        sm = np.ascontiguousarray(shadow_mask, np.uint32)
        sd = np.ascontiguousarray(shadow_distance, np.float32)
        dd = np.ascontiguousarray(draw_distance, np.float32)
        iv = np.ascontiguousarray(in_view, np.uint8)
        n = self.lib.ref_process_shadow_casters_text(self.h, len(ids), _p(ids, _i), _p(m, _f), _p(sh, _f), near_z, far_z, light_type, light_mask,
                                                     _p(sm, _u32), _p(sd, _f), _p(dd, _f), _p(iv, _u8), _p(out, _i))
        return out[:n].copy()

    def run_view_serial(self, scene, sorted_occluders, max_triangles, flags=0xFF, view_mask=0xFFFFFFFF):
        """DrawOccluders' NON-threaded branch over a sorted occluder list, then query + predicate.
        Returns (visible ids, (numTriangles_, query size, visible, activeOccluders_))."""
        out = np.zeros(max(self.n, 1), np.int32)
        counts = np.zeros(4, np.int32)
        sel = np.ascontiguousarray(sorted_occluders, np.int32)
        aabb = np.ascontiguousarray(scene.occ_aabb, np.float32)
        model = np.ascontiguousarray(scene.occ_model, np.float32)
        vtx = np.ascontiguousarray(scene.mesh_vtx)
        idx = np.ascontiguousarray(scene.mesh_idx)
        n = self.lib.ref_run_view_serial(self.h, len(sel), _p(sel, _i), _p(aabb, _f), _p(model, _f), vtx.ctypes.data, scene.mesh_stride,
                                         idx.ctypes.data, idx.itemsize, idx.shape[0], scene.cull_mode, int(max_triangles), flags, view_mask,
                                         _p(out, _i), out.shape[0], _p(counts, _i))
        assert n >= 0, "the reference buffer was created threaded"
        return out[:n].copy(), counts

    def raycast(self, origin, direction, max_distance=np.inf, flags=0xFF, view_mask=0xFFFFFFFF, single=False):
        """Octree::Raycast / RaycastSingle, RAY_AABB level.  Returns (ids, distances, positions) in result_ order."""
        ray = np.array(list(origin) + list(direction) + [max_distance], np.float32)
        cap = max(self.n, 1)
        ids = np.zeros(cap, np.int32)
        dist = np.zeros(cap, np.float32)
        pos = np.zeros((cap, 3), np.float32)
        n = self.lib.ref_raycast(self.h, _p(ray, _f), flags, view_mask, int(single), _p(ids, _i), _p(dist, _f), _p(pos, _f), cap)
        return ids[:n].copy(), dist[:n].copy(), pos[:n].copy()

    def sort_by_key(self, keys, vals):
        """The reference's Sort (Container/Sort.h) with a key less-than compare; returns (sorted keys, payload in sorted order)."""
        k = np.ascontiguousarray(keys, np.float32).copy()
        v = np.ascontiguousarray(vals, np.int32).copy()
        self.lib.ref_sort_by_key(len(k), _p(k, _f), _p(v, _i))
        return k, v

    def check_visibility(self, use_buffer=True):
        out = np.zeros(max(self.n, 1), np.int32)
        n = self.lib.ref_check_visibility(self.h, int(use_buffer), _p(out, _i), out.shape[0])
        return out[:n].copy()

    def run_view(self, scene, flags=0xFF, view_mask=0xFFFFFFFF):
        """Whole view, threaded-branch semantics (View.cpp:2258-2273, 873-878, 885-921).  View must be set already."""
        out = np.zeros(max(self.n, 1), np.int32)
        counts = np.zeros(3, np.int32)
        times = np.zeros(4, np.float64)
        n = self.lib.ref_run_view(self.h, scene.k, _p(scene.occ_aabb, _f), _p(scene.occ_model, _f), scene.mesh_vtx.ctypes.data,
                                  scene.mesh_stride, scene.mesh_idx.ctypes.data, scene.mesh_idx.dtype.itemsize, scene.mesh_idx.shape[0],
                                  scene.cull_mode, flags, view_mask, _p(out, _i), out.shape[0], _p(counts, _i),
                                  times.ctypes.data_as(C.POINTER(C.c_double)))
        return out[:n].copy(), counts, times

    def check_in_view(self, view, cam_pos, ortho=False, draw_dist=None, use_buffer=True):
        """View.cpp:177-211 over the last query() result.  Returns (ids in view, (minZ, maxZ))."""
        out = np.zeros(max(self.n, 1), np.int32)
        mm = np.zeros(2, np.float32)
        v12 = np.ascontiguousarray(view, dtype=np.float32)
        cp = np.ascontiguousarray(cam_pos, dtype=np.float32)
        dd = None if draw_dist is None else np.ascontiguousarray(draw_dist, dtype=np.float32)
        n = self.lib.ref_check_in_view(self.h, int(use_buffer), _p(v12, _f), _p(cp, _f), int(ortho), _p(dd, _f), _p(out, _i), out.shape[0], _p(mm, _f))
        return out[:n].copy(), mm

    def check_in_view_text(self, draw_dist=None, use_buffer=True):
        """CheckVisibilityWork compiled from the reference's own text (View.cpp:160-227) over the last query() result, the camera last given
        to camera_set() being the cull camera.  Returns (ids in view, (minZ, maxZ))."""
        out = np.zeros(max(self.n, 1), np.int32)
        mm = np.zeros(2, np.float32)
        dd = None if draw_dist is None else np.ascontiguousarray(draw_dist, dtype=np.float32)
        n = self.lib.ref_check_in_view_text(self.h, int(use_buffer), _p(dd, _f), _p(out, _i), out.shape[0], _p(mm, _f))
        return out[:n].copy(), mm

    def num_threads(self):
        return self.lib.ref_num_threads(self.h)
