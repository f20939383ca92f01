"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the C restatement in u3d_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)
_u8 = C.POINTER(C.c_ubyte)
_u32 = C.POINTER(C.c_uint)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class _Tree(C.Structure):
    _fields_ = [("m", C.c_int), ("parent", _i), ("box6", _f), ("first", _i), ("count", _i), ("n", C.c_int), ("aabb6", _f),
                ("flags", _u8), ("view_mask", _u32), ("occludee", _u8), ("cast_shadows", _u8)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "u3d_oracle.c")):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_ob_create.restype = C.c_void_p
        L.orc_ob_destroy.argtypes = [C.c_void_p]
        L.orc_ob_set_size.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for n in ("orc_ob_width", "orc_ob_height", "orc_ob_num_levels", "orc_ob_cull_mode"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_ob_num_triangles.argtypes = [C.c_void_p]
        L.orc_ob_num_triangles.restype = C.c_uint
        L.orc_ob_ref_undefined.argtypes = [C.c_void_p]
        L.orc_ob_ref_undefined.restype = C.c_ulonglong
        L.orc_ob_set_view.argtypes = [C.c_void_p, _f, _f, C.c_int]
        L.orc_ob_get_viewproj.argtypes = [C.c_void_p, _f]
        L.orc_ob_set_max_triangles.argtypes = [C.c_void_p, C.c_uint]
        L.orc_ob_set_cull_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_ob_reset.argtypes = [C.c_void_p]
        L.orc_ob_clear.argtypes = [C.c_void_p]
        L.orc_ob_draw_batch.argtypes = [C.c_void_p, _f, C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
        L.orc_ob_build_hierarchy.argtypes = [C.c_void_p]
        L.orc_ob_is_visible.argtypes = [C.c_void_p, _f]
        L.orc_ob_is_visible_many.argtypes = [C.c_void_p, C.c_int, _f, _u8]
        L.orc_ob_read_depth.argtypes = [C.c_void_p, _i]
        L.orc_ob_read_mip.argtypes = [C.c_void_p, C.c_int, _i]
        L.orc_ob_mip_size.argtypes = [C.c_void_p, C.c_int, _i, _i]
        L.orc_frustum_test.argtypes = [_f, _f, C.c_int]
        L.orc_query.argtypes = [C.POINTER(_Tree), C.c_int, _f, C.c_void_p, C.c_uint, C.c_uint, _i]
        L.orc_check_visibility.argtypes = [C.POINTER(_Tree), C.c_void_p, _i, C.c_int, _i]
        L.orc_check_in_view.argtypes = [C.POINTER(_Tree), C.c_void_p, _f, _f, C.c_int, _f, _i, C.c_int, _i, _f]
        _lib = L
    return _lib


class OracleOB:
    """C restatement of OcclusionBuffer (OcclusionBuffer.h:101-158)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_ob_create()

    def close(self):
        if self.h:
            self.L.orc_ob_destroy(self.h)
            self.h = None

    def set_size(self, w, h):
        return bool(self.L.orc_ob_set_size(self.h, w, h))

    def size(self):
        return self.L.orc_ob_width(self.h), self.L.orc_ob_height(self.h)

    def set_view(self, v):
        self.L.orc_ob_set_view(self.h, _p(v.view, _f), _p(v.proj, _f), int(v.reverse_culling))

    def viewproj(self):
        out = np.zeros((4, 4), np.float32)
        self.L.orc_ob_get_viewproj(self.h, _p(out, _f))
        return out

    def set_max_triangles(self, n):
        self.L.orc_ob_set_max_triangles(self.h, n)

    def set_cull_mode(self, m):
        self.L.orc_ob_set_cull_mode(self.h, m)

    def cull_mode(self):
        return self.L.orc_ob_cull_mode(self.h)

    def clear(self):
        self.L.orc_ob_clear(self.h)

    def reset(self):
        self.L.orc_ob_reset(self.h)

    def draw_batch(self, model12, vtx, stride, idx=None, start=0, count=0):
        model12 = np.ascontiguousarray(model12, dtype=np.float32)
        return bool(self.L.orc_ob_draw_batch(self.h, _p(model12, _f), vtx.ctypes.data, stride, None if idx is None else idx.ctypes.data,
                                             0 if idx is None else idx.dtype.itemsize, start, count))

    def build_hierarchy(self):
        self.L.orc_ob_build_hierarchy(self.h)

    def num_triangles(self):
        return self.L.orc_ob_num_triangles(self.h)

    def ref_undefined(self):
        """How often the geometry drawn so far (since creation) would have made the reference write outside its padded allocation."""
        return int(self.L.orc_ob_ref_undefined(self.h))

    def is_visible(self, aabb6):
        b = np.ascontiguousarray(aabb6, dtype=np.float32)
        return bool(self.L.orc_ob_is_visible(self.h, _p(b, _f)))

    def is_visible_many(self, aabb6):
        b = np.ascontiguousarray(aabb6, dtype=np.float32)
        out = np.zeros(b.shape[0], np.uint8)
        self.L.orc_ob_is_visible_many(self.h, b.shape[0], _p(b, _f), _p(out, _u8))
        return out

    def depth(self):
        w, h = self.size()
        out = np.zeros((h, w), np.int32)
        self.L.orc_ob_read_depth(self.h, _p(out, _i))
        return out

    def mips(self):
        res = []
        for lv in range(self.L.orc_ob_num_levels(self.h)):
            w, h = C.c_int(), C.c_int()
            self.L.orc_ob_mip_size(self.h, lv, C.byref(w), C.byref(h))
            out = np.zeros((h.value, w.value, 2), np.int32)
            self.L.orc_ob_read_mip(self.h, lv, _p(out, _i))
            res.append(out)
        return res

    def draw_scene_view_serial(self, scene, view, sorted_occluders, max_triangles):
        """View::DrawOccluders, non-threaded branch (View.cpp:2236-2256) over a sorted occluder list: IsVisible gate after the first occluder,
        AddTriangles + DrawTriangles per occluder, stop after the first failed budget check.  Returns (activeOccluders, numTriangles_)."""
        self.set_view(view)
        self.set_max_triangles(int(max_triangles))
        self.clear()
        active = 0
        for j, i in enumerate(sorted_occluders):
            if j > 0 and not self.is_visible(scene.occ_aabb[i]):
                continue
            active += 1
            self.set_cull_mode(scene.cull_mode)
            ok = self.draw_batch(scene.occ_model[i], scene.mesh_vtx, scene.mesh_stride, scene.mesh_idx, 0, scene.mesh_idx.shape[0])
            if not ok:
                break
        self.build_hierarchy()
        return active, self.num_triangles()

    def draw_scene_view_selected(self, scene, view, selected, max_triangles):
        """DrawOccluders' threaded branch over an explicit (sorted, budget-trimmed) occluder list.  Returns submitted triangles."""
        self.set_view(view)
        self.set_max_triangles(int(max_triangles))
        self.clear()
        sub = 0
        for i in selected:
            self.set_cull_mode(scene.cull_mode)
            self.draw_batch(scene.occ_model[i], scene.mesh_vtx, scene.mesh_stride, scene.mesh_idx, 0, scene.mesh_idx.shape[0])
            sub += scene.mesh_idx.shape[0] // 3
        self.build_hierarchy()
        return sub

    def draw_scene_view(self, scene, view):
        """Threaded-branch view set-up (View.cpp:2258-2273): select occluders by IsInsideFast in scene order, draw all, build mips.
        Returns number of submitted triangles."""
        self.set_view(view)
        self.set_max_triangles(1 << 30)
        self.clear()
        sub = 0
        for i in range(scene.k):
            if frustum_test(view.planes, scene.occ_aabb[i], fast=True) == 0:
                continue
            self.set_cull_mode(scene.cull_mode)
            self.draw_batch(scene.occ_model[i], scene.mesh_vtx, scene.mesh_stride, scene.mesh_idx, 0, scene.mesh_idx.shape[0])
            sub += scene.mesh_idx.shape[0] // 3
        self.build_hierarchy()
        return sub


def frustum_test(planes, aabb6, fast=False):
    p = np.ascontiguousarray(planes, dtype=np.float32)
    b = np.ascontiguousarray(aabb6, dtype=np.float32)
    return lib().orc_frustum_test(_p(p, _f), _p(b, _f), int(fast))


class OracleTree:
    """Flattened DFS octree (see OrcTree in u3d_oracle.c).  `flat` = dict(parent, box, first, count, order) and the per-drawable
    arrays are given in ORIGINAL drawable order; they are permuted to DFS order here."""

    def __init__(self, flat, aabb6, flags, view_mask, occludee, cast_shadows):
        self.order = np.ascontiguousarray(flat["order"], dtype=np.int32)
        o = self.order
        self.parent = np.ascontiguousarray(flat["parent"], dtype=np.int32)
        self.box = np.ascontiguousarray(flat["box"], dtype=np.float32)
        self.first = np.ascontiguousarray(flat["first"], dtype=np.int32)
        self.count = np.ascontiguousarray(flat["count"], dtype=np.int32)
        self.aabb = np.ascontiguousarray(aabb6[o], dtype=np.float32)
        self.flags = np.ascontiguousarray(flags[o], dtype=np.uint8)
        self.view_mask = np.ascontiguousarray(view_mask[o], dtype=np.uint32)
        self.occludee = np.ascontiguousarray(occludee[o], dtype=np.uint8)
        self.cast_shadows = np.ascontiguousarray(cast_shadows[o], dtype=np.uint8)
        self.t = _Tree(len(self.parent), _p(self.parent, _i), _p(self.box, _f), _p(self.first, _i), _p(self.count, _i), len(o),
                       _p(self.aabb, _f), _p(self.flags, _u8), _p(self.view_mask, _u32), _p(self.occludee, _u8), _p(self.cast_shadows, _u8))

    def query(self, mode, planes, ob=None, flags=0xFF, view_mask=0xFFFFFFFF, as_ids=True):
        p = np.zeros(24, np.float32)
        flat_p = np.asarray(planes, np.float32).reshape(-1)
        p[:flat_p.shape[0]] = flat_p
        out = np.zeros(max(len(self.order), 1), np.int32)
        n = lib().orc_query(C.byref(self.t), mode, _p(p, _f), ob.h if ob is not None else None, flags, view_mask, _p(out, _i))
        pos = out[:n].copy()
        return self.order[pos] if as_ids else pos

    def query_zone_occluder(self, planes, is_occluder, view_mask=0xFFFFFFFF):
        """ZoneOccluderOctreeQuery (View.cpp:87-113).  is_occluder is per drawable id.  Returns drawable ids in result_ order."""
        p = np.ascontiguousarray(planes, np.float32).reshape(-1)
        occ = np.ascontiguousarray(np.asarray(is_occluder, np.uint8)[self.order])
        out = np.zeros(max(len(self.order), 1), np.int32)
        L = lib()
        L.orc_query_zone_occluder.argtypes = [C.c_void_p, _f, _u8, C.c_uint, _i]
        n = L.orc_query_zone_occluder(C.byref(self.t), _p(p, _f), _p(occ, _u8), view_mask, _p(out, _i))
        return self.order[out[:n]]

    def raycast(self, origin, direction, max_distance=np.inf, flags=0xFF, view_mask=0xFFFFFFFF, single=False):
        """Octree::Raycast / RaycastSingle at RAY_AABB level.  Returns (drawable ids, distances) in result_ order."""
        ray = np.array(list(origin) + list(direction) + [max_distance], np.float32)
        pos = np.zeros(max(len(self.order), 1), np.int32)
        dist = np.zeros(max(len(self.order), 1), np.float32)
        L = lib()
        L.orc_raycast.argtypes = [C.c_void_p, _f, C.c_uint, C.c_uint, C.c_int, _i, _f]
        n = L.orc_raycast(C.byref(self.t), _p(ray, _f), flags, view_mask, int(single), _p(pos, _i), _p(dist, _f))
        return self.order[pos[:n]], dist[:n].copy()

    def check_visibility(self, ob, cand_pos):
        cand = np.ascontiguousarray(cand_pos, dtype=np.int32)
        out = np.zeros(max(len(cand), 1), np.int32)
        n = lib().orc_check_visibility(C.byref(self.t), ob.h if ob is not None else None, _p(cand, _i), len(cand), _p(out, _i))
        return out[:n].copy()

    def check_in_view(self, ob, cand_pos, view, cam_pos, ortho=False, draw_dist=None):
        """View.cpp:177-211.  draw_dist is per drawable id (original order).  Returns (DFS positions in view, (minZ, maxZ))."""
        cand = np.ascontiguousarray(cand_pos, dtype=np.int32)
        out = np.zeros(max(len(cand), 1), np.int32)
        mm = np.zeros(2, np.float32)
        v12 = np.ascontiguousarray(view, dtype=np.float32)
        cp = np.ascontiguousarray(cam_pos, dtype=np.float32)
        dd = None if draw_dist is None else np.ascontiguousarray(np.asarray(draw_dist, np.float32)[self.order])
        n = lib().orc_check_in_view(C.byref(self.t), ob.h if ob is not None else None, _p(v12, _f), _p(cp, _f), int(ortho), _p(dd, _f), _p(cand, _i),
                                    len(cand), _p(out, _i), _p(mm, _f))
        return out[:n].copy(), mm


def select_occluders(scene, view, size_threshold, max_triangles, draw_dist=None, num_triangles=None):
    """View::UpdateOccluders + DrawOccluders' budget.  Returns (sorted occluder indices, sort values, number submitted under the budget)."""
    k = scene.k
    tris = np.full(k, scene.mesh_idx.shape[0] // 3, np.uint32) if num_triangles is None else np.ascontiguousarray(num_triangles, np.uint32)
    dd = None if draw_dist is None else np.ascontiguousarray(draw_dist, np.float32)
    aabb = np.ascontiguousarray(scene.occ_aabb, np.float32)
    planes = np.ascontiguousarray(view.planes, np.float32)
    v12 = np.ascontiguousarray(view.view, np.float32)
    cp = np.ascontiguousarray(view.position, np.float32)
    out = np.zeros(max(k, 1), np.int32)
    keys = np.zeros(max(k, 1), np.float32)
    sub = C.c_int()
    L = lib()
    L.orc_select_occluders.argtypes = [C.c_int, _f, C.POINTER(C.c_uint), _f, _f, _f, _f, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint,
                                       _i, _f, C.POINTER(C.c_int)]
    n = L.orc_select_occluders(k, _p(aabb, _f), tris.ctypes.data_as(C.POINTER(C.c_uint)), _p(dd, _f), _p(planes, _f), _p(v12, _f), _p(cp, _f),
                               int(view.ortho), view.half_view_size, view.ortho_size, view.far, size_threshold, int(max_triangles), _p(out, _i),
                               _p(keys, _f), C.byref(sub))
    return out[:n].copy(), keys[:n].copy(), sub.value


def process_shadow_casters(ids, aabb6, cast_shadows, split, light_type, light_mask, shadow_mask, shadow_distance, draw_distance, in_view, main_view,
                           main_pos, main_ortho=False):
    """View::ProcessShadowCasters' per-drawable filter (View.cpp:2409-2489) over `ids` (arrays indexed by drawable id); returns the kept ids."""
    ids = np.ascontiguousarray(ids, np.int32)
    out = np.zeros(max(len(ids), 1), np.int32)
    bb = np.ascontiguousarray(aabb6, np.float32)
    cs = np.ascontiguousarray(cast_shadows, np.uint8)
    sm = np.ascontiguousarray(shadow_mask, np.uint32)
    sd = np.ascontiguousarray(shadow_distance, np.float32)
    dd = np.ascontiguousarray(draw_distance, np.float32)
    iv = np.ascontiguousarray(in_view, np.uint8)
    mv = np.ascontiguousarray(main_view, np.float32).reshape(-1)
    mp = np.ascontiguousarray(main_pos, np.float32)
    L = lib()
    L.orc_process_shadow_casters.argtypes = [C.c_int, _i, _f, _u8, _f, _f, _f, C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint,
                                             C.POINTER(C.c_uint), _f, _f, _u8, _f, _f, C.c_int, _i]
    n = L.orc_process_shadow_casters(len(ids), _p(ids, _i), _p(bb, _f), _p(cs, _u8), _p(split["light_view"], _f), _p(split["shadow_planes"], _f),
                                     _p(split["lvf_planes"], _f), split["lvf_max_z"], split["shadow_far"], int(split["shadow_ortho"]), light_type,
                                     light_mask, sm.ctypes.data_as(C.POINTER(C.c_uint)), _p(sd, _f), _p(dd, _f), _p(iv, _u8), _p(mv, _f), _p(mp, _f),
                                     int(main_ortho), _p(out, _i))
    return out[:n].copy()


def sort_by_key(keys, vals):
    """The restated Sort of Container/Sort.h over (key, payload) pairs."""
    k = np.ascontiguousarray(keys, np.float32).copy()
    v = np.ascontiguousarray(vals, np.int32).copy()
    L = lib()
    L.orc_sort_by_key.argtypes = [C.c_int, _f, _i]
    L.orc_sort_by_key(len(k), _p(k, _f), _p(v, _i))
    return k, v


def ray_hit_distance(ray6, box6):
    L = lib()
    L.orc_ray_hit_distance.argtypes = [_f, _f]
    L.orc_ray_hit_distance.restype = C.c_float
    r = np.ascontiguousarray(ray6, np.float32)
    b = np.ascontiguousarray(box6, np.float32)
    return float(L.orc_ray_hit_distance(_p(r, _f), _p(b, _f)))
