"""ctypes binding of libu3dgpu.so (include/u3d_gpu.h) plus thin Python mirrors of the reference-facing objects.

The library is the product; this file is plumbing for tests, smoke and bench.  It never falls back to a CPU
implementation: if the shared library is missing it raises, and every compute call goes to the CUDA kernels.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("U3D_LIB") or os.path.join(HERE, "lib", "libu3dgpu.so")  # U3D_LIB: build variants in tuning sweeps

SHADOW_SPLIT_DTYPE = np.dtype([("light_view", np.float32, 12), ("shadow_camera_frustum", np.float32, 24), ("light_view_frustum", np.float32, 24),
                               ("light_view_frustum_max_z", np.float32), ("shadow_far_clip", np.float32), ("shadow_orthographic", np.int32),
                               ("light_type", np.int32), ("light_mask", np.uint32), ("skip", np.int32), ("reserved", np.int32, 2)])
RAY_HIT_DTYPE = np.dtype([("position", np.float32, 3), ("normal", np.float32, 3), ("distance", np.float32), ("drawable", np.uint32)])
QUERY_FRUSTUM, QUERY_OCCLUDED, QUERY_SHADOWCASTER, QUERY_SPHERE, QUERY_BOX, QUERY_POINT, QUERY_RAY, QUERY_RAY_CANDIDATES, QUERY_ZONE_OCCLUDER = 0, 1, 2, 3, 4, 5, 6, 7, 8
STAGE_QUERY, STAGE_VISIBLE, STAGE_INVIEW = 0, 1, 2
CULL_NONE, CULL_CCW, CULL_CW = 0, 1, 2
ERR_CAPACITY = 4

_f = C.POINTER(C.c_float)
_i32 = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)
_u32 = C.POINTER(C.c_uint32)
_vp = C.c_void_p


class U3DView(C.Structure):
    """struct u3d_view of include/u3d_gpu.h."""
    _fields_ = [("view", C.c_float * 12), ("projection", C.c_float * 16), ("planes", C.c_float * 24), ("near_clip", C.c_float),
                ("far_clip", C.c_float), ("drawable_flags", C.c_uint32), ("view_mask", C.c_uint32), ("reverse_culling", C.c_int32),
                ("query_mode", C.c_int32), ("use_occlusion", C.c_int32), ("orthographic", C.c_int32), ("camera_position", C.c_float * 3),
                ("reserved", C.c_int32)]


VIEW_DTYPE = np.dtype([("view", np.float32, 12), ("projection", np.float32, 16), ("planes", np.float32, 24), ("near_clip", np.float32),
                       ("far_clip", np.float32), ("drawable_flags", np.uint32), ("view_mask", np.uint32), ("reverse_culling", np.int32),
                       ("query_mode", np.int32), ("use_occlusion", np.int32), ("orthographic", np.int32), ("camera_position", np.float32, 3),
                       ("reserved", np.int32)])
assert VIEW_DTYPE.itemsize == C.sizeof(U3DView) == 256

# name -> (restype, argtypes); the list doubles as the symbol inventory checked by tests/test_capi_symbols.py
SIGNATURES = {
    "u3d_last_error": (C.c_char_p, []),
    "u3d_abi_version": (C.c_int, []),
    "u3d_debug_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "u3d_debug_read_cull_prof": (C.c_int, [C.POINTER(C.c_uint64)]),
    "u3d_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "u3d_ctx_destroy": (None, [_vp]),
    "u3d_ctx_device": (C.c_int, [_vp]),
    "u3d_ctx_synchronize": (C.c_int, [_vp, _vp]),
    "u3d_octree_create": (C.c_int, [_f, _f, C.c_uint32, C.POINTER(_vp)]),
    "u3d_octree_destroy": (None, [_vp]),
    "u3d_octree_add": (C.c_int, [_vp, C.c_uint32, _f, _u8, _u32, _u8, _u8, _u32]),
    "u3d_octree_set_draw_distances": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _f]),
    "u3d_octree_update": (C.c_int, [_vp, C.c_uint32, _f]),
    "u3d_octree_remove": (C.c_int, [_vp, C.c_uint32]),
    "u3d_octree_counts": (C.c_int, [_vp, _u32, _u32]),
    "u3d_octree_flatten": (C.c_int, [_vp, _i32, _i32, _f, _i32, _i32, _i32]),
    "u3d_scene_create": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "u3d_scene_create_flat": (C.c_int, [_vp, C.c_uint32, _i32, _i32, _f, _i32, _i32, C.c_uint32, _i32, _f, _u8, _u32, _u8, _u8, C.POINTER(_vp)]),
    "u3d_scene_set_draw_distances": (C.c_int, [_vp, C.c_uint32, _f]),
    "u3d_scene_update_boxes": (C.c_int, [_vp, C.c_uint32, _u32, _f, _vp]),
    "u3d_scene_destroy": (None, [_vp]),
    "u3d_scene_counts": (C.c_int, [_vp, _u32, _u32]),
    "u3d_mesh_create": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "u3d_mesh_destroy": (None, [_vp]),
    "u3d_occluders_create": (C.c_int, [_vp, C.c_uint32, _f, _f, C.POINTER(_vp), C.c_uint32, _u32, _u32, _u32, C.c_int, C.POINTER(_vp)]),
    "u3d_occluders_destroy": (None, [_vp]),
    "u3d_ob_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "u3d_ob_destroy": (None, [_vp]),
    "u3d_ob_set_size": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "u3d_ob_set_view": (C.c_int, [_vp, _f, _f, C.c_float, C.c_float, C.c_int]),
    "u3d_ob_set_max_triangles": (C.c_int, [_vp, C.c_uint32]),
    "u3d_ob_set_cull_mode": (C.c_int, [_vp, C.c_int]),
    "u3d_ob_reset": (C.c_int, [_vp]),
    "u3d_ob_clear": (C.c_int, [_vp]),
    "u3d_ob_add_triangles": (C.c_int, [_vp, _f, _vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "u3d_ob_add_triangles_indexed": (C.c_int, [_vp, _f, _vp, C.c_uint32, _vp, C.c_uint32, C.c_uint32, C.c_uint32]),
    "u3d_ob_add_triangles_mesh": (C.c_int, [_vp, _f, _vp, C.c_uint32, C.c_uint32]),
    "u3d_ob_draw_triangles": (C.c_int, [_vp]),
    "u3d_ob_build_depth_hierarchy": (C.c_int, [_vp]),
    "u3d_ob_is_visible": (C.c_int, [_vp, _f]),
    "u3d_ob_is_visible_many": (C.c_int, [_vp, _f, C.c_uint32, _u8]),
    "u3d_ob_get_buffer": (C.c_int, [_vp, _i32]),
    "u3d_ob_get_mip": (C.c_int, [_vp, C.c_int, _i32]),
    "u3d_ob_num_levels": (C.c_int, [_vp]),
    "u3d_ob_mip_size": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "u3d_ob_get_width": (C.c_int, [_vp]),
    "u3d_ob_get_height": (C.c_int, [_vp]),
    "u3d_ob_get_num_triangles": (C.c_uint32, [_vp]),
    "u3d_ob_get_max_triangles": (C.c_uint32, [_vp]),
    "u3d_ob_get_cull_mode": (C.c_int, [_vp]),
    "u3d_ob_is_threaded": (C.c_int, [_vp]),
    "u3d_ob_get_view_proj": (C.c_int, [_vp, _f]),
    "u3d_octree_query": (C.c_int, [_vp, _vp, _f, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u32, C.c_uint32, _u32, _u32]),
    "u3d_batch_create": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(_vp)]),
    "u3d_batch_destroy": (None, [_vp]),
    "u3d_batch_set_scene": (C.c_int, [_vp, _vp]),
    "u3d_batch_set_views": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "u3d_batch_run": (C.c_int, [_vp, C.c_int, _vp]),
    "u3d_batch_run_part": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "u3d_batch_run_timed": (C.c_int, [_vp, C.c_int, _vp, _f]),
    "u3d_batch_read_counts": (C.c_int, [_vp, _u32, _vp]),
    "u3d_batch_read_packed": (C.c_int, [_vp, _u32, _u32, C.c_uint64, _vp]),
    "u3d_batch_read_z_ranges": (C.c_int, [_vp, _f, _vp]),
    "u3d_batch_read_view": (C.c_int, [_vp, C.c_uint32, _u32, C.c_uint32, _u32, _vp]),
    "u3d_batch_read_depth": (C.c_int, [_vp, C.c_uint32, _i32, _vp]),
    "u3d_batch_read_mip": (C.c_int, [_vp, C.c_uint32, C.c_int, _i32, _vp]),
    "u3d_batch_read_tri_stats": (C.c_int, [_vp, _u32, _vp]),
    "u3d_batch_read_setup_records": (C.c_int, [_vp, C.c_uint32, _i32, C.c_uint32, _u32, _vp]),
    "u3d_batch_cull_views": (C.c_int, [_vp, _vp, C.c_uint32, C.c_int, _u32, _u32, _u32, C.c_uint64, _vp]),
    "u3d_batch_device_results": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), _u32]),
    "u3d_batch_pack_device": (C.c_int, [_vp, _vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "u3d_batch_last_launches": (C.c_uint32, [_vp]),
    "u3d_batch_set_pipelined": (C.c_int, [_vp, C.c_int]),
    "u3d_octree_set_occluder_flags": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint8)]),
    "u3d_scene_set_occluder_flags": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_uint8)]),
    "u3d_scene_set_shadow_params": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_float)]),
    "u3d_batch_process_shadow_casters": (C.c_int, [_vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint8), C.c_uint32, _vp]),
    "u3d_scene_raycast": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_float), C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_uint32), _vp, _vp]),
    "u3d_sort_by_key": (C.c_int, [C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "u3d_balance_views": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_uint32)]),
    "u3d_batch_set_occluder_policy": (C.c_int, [_vp, C.c_int, C.c_float, C.c_uint32, C.POINTER(C.c_float), _vp]),
    "u3d_occluders_set_draw_distances": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_float)]),
    "u3d_batch_read_active_occluders": (C.c_int, [_vp, C.POINTER(C.c_uint32), _vp]),
    "u3d_batch_read_occluder_selection": (C.c_int, [_vp, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.c_uint32), _vp]),
    "u3d_gather_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(_vp)]),
    "u3d_gather_destroy": (None, [_vp]),
    "u3d_gather_export": (C.c_int, [_vp, _vp]),
    "u3d_gather_connect": (C.c_int, [_vp, _vp]),
    "u3d_gather_connect_local": (C.c_int, [_vp, C.POINTER(_vp)]),
    "u3d_batch_push_results": (C.c_int, [_vp, _vp, _vp]),
    "u3d_gather_wait": (C.c_int, [_vp, _vp]),
    "u3d_gather_release": (C.c_int, [_vp, _vp]),
    "u3d_gather_buffers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "u3d_gather_read": (C.c_int, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _vp]),
    "u3d_gather_status": (C.c_int, [_vp, _vp]),
}

_lib = None


class U3DError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("u3d error %d: %s" % (code, msg))
        self.code = code


ABI_VERSION = 2  # include/u3d_gpu.h: U3D_ABI_VERSION


def lib():
    """Load libu3dgpu.so.  Fails loudly when the CUDA extension has not been built: there is no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libu3dgpu.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "or `make -C urho3d_b200/csrc`; there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.u3d_abi_version() != ABI_VERSION:
            raise RuntimeError("libu3dgpu.so reports ABI version %d, this binding was written for %d: rebuild it (make -C urho3d_b200/csrc)"
                               % (L.u3d_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def sort_by_key(keys, values):
    """The reference's Sort procedure (Container/Sort.h) over float keys with a uint32 payload; host code."""
    k = np.ascontiguousarray(keys, np.float32).copy()
    v = np.ascontiguousarray(values, np.uint32).copy()
    _chk(lib().u3d_sort_by_key(len(k), _p(k, _f), _p(v, _u32)))
    return k, v


def balance_views(costs, world):
    """u3d_balance_views: per-view cost columns [n] or [n, D] -> owner rank of every view (uint32 [n]); host code."""
    c = np.ascontiguousarray(costs, np.float64)
    if c.ndim == 1:
        c = c[:, None]
    owner = np.zeros(c.shape[0], np.uint32)
    _chk(lib().u3d_balance_views(c.shape[0], c.shape[1], c.ctypes.data_as(C.POINTER(C.c_double)), world, _p(owner, _u32)))
    return owner


def debug_set_option(name, value):
    _chk(lib().u3d_debug_set_option(name.encode(), int(value)))


def read_cull_prof():
    out = np.zeros(8, np.uint64)
    _chk(lib().u3d_debug_read_cull_prof(out.ctypes.data_as(C.POINTER(C.c_uint64))))
    return out


def _chk(rc):
    if rc < 0:
        raise U3DError(-rc, lib().u3d_last_error().decode())
    return rc


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _arr(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Context:
    def __init__(self, device=0):
        h = _vp()
        _chk(lib().u3d_ctx_create(device, C.byref(h)))
        self.h = h

    def synchronize(self, stream=None):
        _chk(lib().u3d_ctx_synchronize(self.h, stream))

    def close(self):
        if self.h:
            lib().u3d_ctx_destroy(self.h)
            self.h = None


class Octree:
    """Host octree mirror (no GPU needed).  Same structure as the reference's Octree for the same inserts."""

    def __init__(self, mn=(-1000.0,) * 3, mx=(1000.0,) * 3, levels=8):
        h = _vp()
        mn = np.asarray(mn, np.float32)
        mx = np.asarray(mx, np.float32)
        _chk(lib().u3d_octree_create(_p(mn, _f), _p(mx, _f), levels, C.byref(h)))
        self.h = h

    def add(self, aabb6, flags=None, view_mask=None, occludee=None, cast_shadows=None):
        a = _arr(aabb6, np.float32).reshape(-1, 6)
        fl, vm, oc, cs = _arr(flags, np.uint8), _arr(view_mask, np.uint32), _arr(occludee, np.uint8), _arr(cast_shadows, np.uint8)
        first = C.c_uint32()
        _chk(lib().u3d_octree_add(self.h, a.shape[0], _p(a, _f), _p(fl, _u8), _p(vm, _u32), _p(oc, _u8), _p(cs, _u8), C.byref(first)))
        return first.value

    def update(self, idx, aabb6):
        """Returns True when the drawable stayed in its octant (the flattened layout is unchanged)."""
        a = _arr(aabb6, np.float32)
        return bool(_chk(lib().u3d_octree_update(self.h, idx, _p(a, _f))))

    def set_occluder_flags(self, is_occluder, first_id=0):
        f = np.ascontiguousarray(is_occluder, np.uint8)
        _chk(lib().u3d_octree_set_occluder_flags(self.h, first_id, f.shape[0], f.ctypes.data_as(C.POINTER(C.c_uint8))))

    def set_draw_distances(self, draw_distance, first_id=0):
        d = _arr(draw_distance, np.float32)
        _chk(lib().u3d_octree_set_draw_distances(self.h, first_id, d.shape[0], _p(d, _f)))

    def remove(self, idx):
        _chk(lib().u3d_octree_remove(self.h, idx))

    def flatten(self):
        m, n = C.c_uint32(), C.c_uint32()
        _chk(lib().u3d_octree_counts(self.h, C.byref(m), C.byref(n)))
        m, n = m.value, n.value
        parent, level, first, count = (np.zeros(m, np.int32) for _ in range(4))
        box = np.zeros((m, 6), np.float32)
        order = np.zeros(max(n, 1), np.int32)
        _chk(lib().u3d_octree_flatten(self.h, _p(parent, _i32), _p(level, _i32), _p(box, _f), _p(first, _i32), _p(count, _i32), _p(order, _i32)))
        return dict(parent=parent, level=level, box=box, first=first, count=count, order=order[:n])

    def close(self):
        if self.h:
            lib().u3d_octree_destroy(self.h)
            self.h = None


class GpuScene:
    def __init__(self, ctx, octree=None, flat=None, aabb6=None, flags=None, view_mask=None, occludee=None, cast_shadows=None):
        h = _vp()
        if octree is not None:
            _chk(lib().u3d_scene_create(ctx.h, octree.h, C.byref(h)))
        else:
            par, lev = _arr(flat["parent"], np.int32), _arr(flat["level"], np.int32)
            box, fst, cnt = _arr(flat["box"], np.float32), _arr(flat["first"], np.int32), _arr(flat["count"], np.int32)
            order = _arr(flat["order"], np.int32)
            a = _arr(aabb6, np.float32)
            fl, vm, oc, cs = _arr(flags, np.uint8), _arr(view_mask, np.uint32), _arr(occludee, np.uint8), _arr(cast_shadows, np.uint8)
            _chk(lib().u3d_scene_create_flat(ctx.h, len(par), _p(par, _i32), _p(lev, _i32), _p(box, _f), _p(fst, _i32), _p(cnt, _i32), len(order),
                                             _p(order, _i32), _p(a, _f), _p(fl, _u8), _p(vm, _u32), _p(oc, _u8), _p(cs, _u8), C.byref(h)))
        self.h = h
        self.ctx = ctx
        m, n = C.c_uint32(), C.c_uint32()
        _chk(lib().u3d_scene_counts(self.h, C.byref(m), C.byref(n)))
        self.num_octants, self.num_drawables = m.value, n.value

    def set_draw_distances(self, draw_distance):
        d = _arr(draw_distance, np.float32)
        _chk(lib().u3d_scene_set_draw_distances(self.h, d.shape[0], _p(d, _f)))

    def set_shadow_params(self, shadow_mask=None, shadow_distance=None):
        m = None if shadow_mask is None else np.ascontiguousarray(shadow_mask, np.uint32)
        d = None if shadow_distance is None else np.ascontiguousarray(shadow_distance, np.float32)
        n = len(m) if m is not None else (len(d) if d is not None else 0)
        _chk(lib().u3d_scene_set_shadow_params(self.h, n, _p(m, _u32), _p(d, _f)))

    def set_occluder_flags(self, is_occluder):
        f = np.ascontiguousarray(is_occluder, np.uint8)
        _chk(lib().u3d_scene_set_occluder_flags(self.h, f.shape[0], f.ctypes.data_as(C.POINTER(C.c_uint8))))

    def raycast(self, rays7, flags=0xFF, view_mask=0xFFFFFFFF, single=False, capacity=4096, stream=None):
        """Octree::Raycast / RaycastSingle (RAY_AABB) for many rays.  rays7: n x (origin, direction, maxDistance).
        Returns a list of structured arrays (RAY_HIT_DTYPE), one per ray, in result_ order."""
        rays = np.ascontiguousarray(rays7, np.float32).reshape(-1, 7)
        n = rays.shape[0]
        counts = np.zeros(max(n, 1), np.uint32)
        hits = np.zeros((max(n, 1), capacity), RAY_HIT_DTYPE)
        _chk(lib().u3d_scene_raycast(self.h, n, _p(rays, _f), flags, view_mask, int(single), capacity, _p(counts, _u32), hits.ctypes.data, stream))
        return [hits[i, :counts[i]].copy() for i in range(n)]

    def update_boxes(self, ids, aabb6, stream=None):
        """Rewrite the world boxes of drawables that moved inside their octant (no re-flatten)."""
        i = _arr(ids, np.uint32)
        a = _arr(aabb6, np.float32)
        _chk(lib().u3d_scene_update_boxes(self.h, i.shape[0], _p(i, _u32), _p(a, _f), stream))

    def query(self, planes, ob=None, flags=0xFF, view_mask=0xFFFFFFFF, mode=QUERY_FRUSTUM, stage=STAGE_QUERY, capacity=None):
        """Octree::GetDrawables with a frustum-type query (+ optional visibility predicate).  Returns (ids, query_count)."""
        cap = self.num_drawables if capacity is None else capacity
        out = np.zeros(max(cap, 1), np.uint32)
        p = np.zeros(24, np.float32)
        flat_p = np.asarray(planes, np.float32).reshape(-1)
        p[:flat_p.shape[0]] = flat_p  # 6 planes, or the parameters of a sphere / box / point query
        qc, c = C.c_uint32(), C.c_uint32()
        _chk(lib().u3d_octree_query(self.h, ob.h if ob is not None else None, _p(p, _f), flags, view_mask, mode, stage, _p(out, _u32), cap,
                                    C.byref(qc), C.byref(c)))
        return out[:c.value].astype(np.int64), qc.value

    def close(self):
        if self.h:
            lib().u3d_scene_destroy(self.h)
            self.h = None


class Mesh:
    def __init__(self, ctx, vtx, stride, idx=None):
        h = _vp()
        vtx = np.ascontiguousarray(vtx)
        nv = vtx.nbytes // stride
        self._keep = (vtx, idx)
        _chk(lib().u3d_mesh_create(ctx.h, vtx.ctypes.data, stride, nv, None if idx is None else idx.ctypes.data,
                                   0 if idx is None else idx.dtype.itemsize, 0 if idx is None else idx.shape[0], C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().u3d_mesh_destroy(self.h)
            self.h = None


class Occluders:
    def __init__(self, ctx, world_aabb6, model12, meshes, mesh_of=None, start=None, count=None, cull_mode=CULL_CCW):
        h = _vp()
        a, m = _arr(world_aabb6, np.float32), _arr(model12, np.float32)
        n = a.shape[0] if a is not None else 0
        arr = (_vp * max(len(meshes), 1))(*[x.h for x in meshes])
        mo, st, ct = _arr(mesh_of, np.uint32), _arr(start, np.uint32), _arr(count, np.uint32)
        _chk(lib().u3d_occluders_create(ctx.h, n, _p(a, _f), _p(m, _f), arr, len(meshes), _p(mo, _u32), _p(st, _u32), _p(ct, _u32), cull_mode,
                                        C.byref(h)))
        self.h = h
        self.n = n

    def set_draw_distances(self, draw_distance):
        d = np.ascontiguousarray(draw_distance, np.float32)
        _chk(lib().u3d_occluders_set_draw_distances(self.h, d.shape[0], _p(d, _f)))

    def close(self):
        if self.h:
            lib().u3d_occluders_destroy(self.h)
            self.h = None


class OcclusionBuffer:
    """Mirror of class OcclusionBuffer (OcclusionBuffer.h:91-226) over the C-ABI: same method names, argument meaning and returns."""

    def __init__(self, ctx):
        h = _vp()
        _chk(lib().u3d_ob_create(ctx.h, C.byref(h)))
        self.h = h
        self._keep = []

    def SetSize(self, width, height, threaded=False):
        return bool(_chk(lib().u3d_ob_set_size(self.h, width, height, int(threaded))))

    def SetView(self, v):
        """v: urho3d_b200.camera.CameraView (what Camera::GetView/GetProjection/GetNearClip/GetFarClip/GetReverseCulling give)."""
        _chk(lib().u3d_ob_set_view(self.h, _p(v.view, _f), _p(v.proj, _f), v.near, v.far, int(v.reverse_culling)))

    def SetMaxTriangles(self, n):
        _chk(lib().u3d_ob_set_max_triangles(self.h, n))

    def SetCullMode(self, mode):
        _chk(lib().u3d_ob_set_cull_mode(self.h, mode))

    def Reset(self):
        _chk(lib().u3d_ob_reset(self.h))
        self._keep = []

    def Clear(self):
        _chk(lib().u3d_ob_clear(self.h))
        self._keep = []

    def AddTriangles(self, model12, vtx, vtx_size, idx=None, start=0, count=0):
        model12 = _arr(model12, np.float32)
        self._keep += [vtx, idx]
        if idx is None:
            return bool(_chk(lib().u3d_ob_add_triangles(self.h, _p(model12, _f), vtx.ctypes.data, vtx_size, start, count)))
        return bool(_chk(lib().u3d_ob_add_triangles_indexed(self.h, _p(model12, _f), vtx.ctypes.data, vtx_size, idx.ctypes.data,
                                                            idx.dtype.itemsize, start, count)))

    def AddTrianglesMesh(self, model12, mesh, start, count):
        model12 = _arr(model12, np.float32)
        return bool(_chk(lib().u3d_ob_add_triangles_mesh(self.h, _p(model12, _f), mesh.h, start, count)))

    def DrawTriangles(self):
        _chk(lib().u3d_ob_draw_triangles(self.h))
        self._keep = []

    def BuildDepthHierarchy(self):
        _chk(lib().u3d_ob_build_depth_hierarchy(self.h))

    def IsVisible(self, aabb6):
        b = _arr(aabb6, np.float32)
        return bool(_chk(lib().u3d_ob_is_visible(self.h, _p(b, _f))))

    def IsVisibleMany(self, aabb6):
        b = _arr(aabb6, np.float32).reshape(-1, 6)
        out = np.zeros(b.shape[0], np.uint8)
        _chk(lib().u3d_ob_is_visible_many(self.h, _p(b, _f), b.shape[0], _p(out, _u8)))
        return out

    def GetBuffer(self):
        out = np.zeros((self.GetHeight(), self.GetWidth()), np.int32)
        _chk(lib().u3d_ob_get_buffer(self.h, _p(out, _i32)))
        return out

    def GetMips(self):
        res = []
        for lv in range(lib().u3d_ob_num_levels(self.h)):
            w, h = C.c_int(), C.c_int()
            _chk(lib().u3d_ob_mip_size(self.h, lv, C.byref(w), C.byref(h)))
            out = np.zeros((h.value, w.value, 2), np.int32)
            _chk(lib().u3d_ob_get_mip(self.h, lv, _p(out, _i32)))
            res.append(out)
        return res

    def GetWidth(self):
        return lib().u3d_ob_get_width(self.h)

    def GetHeight(self):
        return lib().u3d_ob_get_height(self.h)

    def GetNumTriangles(self):
        return lib().u3d_ob_get_num_triangles(self.h)

    def GetMaxTriangles(self):
        return lib().u3d_ob_get_max_triangles(self.h)

    def GetCullMode(self):
        return lib().u3d_ob_get_cull_mode(self.h)

    def IsThreaded(self):
        return bool(lib().u3d_ob_is_threaded(self.h))

    def GetViewProj(self):
        out = np.zeros((4, 4), np.float32)
        _chk(lib().u3d_ob_get_view_proj(self.h, _p(out, _f)))
        return out

    def close(self):
        if self.h:
            lib().u3d_ob_destroy(self.h)
            self.h = None


def pack_views(views, flags=0xFF, view_mask=0xFFFFFFFF, query_mode=QUERY_OCCLUDED, use_occlusion=True):
    """list of camera.CameraView -> numpy array of struct u3d_view."""
    out = np.zeros(len(views), VIEW_DTYPE)
    for i, v in enumerate(views):
        out[i]["view"] = v.view.reshape(-1)
        out[i]["projection"] = v.proj.reshape(-1)
        out[i]["planes"] = v.planes.reshape(-1)
        out[i]["near_clip"] = v.near
        out[i]["far_clip"] = v.far
        out[i]["reverse_culling"] = int(v.reverse_culling)
        out[i]["orthographic"] = int(getattr(v, "ortho", False))
        out[i]["camera_position"] = getattr(v, "position", np.zeros(3, np.float32))
    out["drawable_flags"] = flags
    out["view_mask"] = view_mask
    out["query_mode"] = query_mode
    out["use_occlusion"] = int(use_occlusion)
    return out


class Batch:
    """The batched many-view entry point."""

    def __init__(self, ctx, scene, occluders, width, height, max_views, out_capacity, tri_pool=0):
        h = _vp()
        _chk(lib().u3d_batch_create(ctx.h, scene.h, occluders.h if occluders is not None else None, width, height, max_views, out_capacity,
                                    tri_pool, C.byref(h)))
        self.h = h
        self.ctx = ctx
        self.width, self.height = width, height + (height & 1)
        self.max_views = max_views
        self.out_capacity = out_capacity
        self.n = 0

    def set_scene(self, scene):
        """Re-point the batch at another GpuScene (after the octree's flattened layout changed)."""
        _chk(lib().u3d_batch_set_scene(self.h, scene.h))

    def set_views(self, packed, stream=None):
        packed = np.ascontiguousarray(packed)
        self._views = packed
        self.n = packed.shape[0]
        _chk(lib().u3d_batch_set_views(self.h, packed.ctypes.data, self.n, stream))

    def run(self, stage=STAGE_VISIBLE, stream=None, part=0):
        _chk(lib().u3d_batch_run_part(self.h, part, stage, stream))

    PHASES = ("clear", "select", "scan", "setup", "fill", "second_pass", "tree", "frustum", "occlusion", "emit")

    def run_timed(self, stage=STAGE_VISIBLE, stream=None):
        """Run with a CUDA event after each phase; returns dict phase -> device milliseconds (synchronises)."""
        ms = np.zeros(len(self.PHASES), np.float32)
        _chk(lib().u3d_batch_run_timed(self.h, stage, stream, _p(ms, _f)))
        return dict(zip(self.PHASES, ms.tolist()))

    def counts(self, stream=None):
        out = np.zeros((self.n, 2), np.uint32)
        _chk(lib().u3d_batch_read_counts(self.h, _p(out, _u32), stream))
        return out

    def packed(self, stream=None, capacity=None):
        cap = self.n * self.out_capacity if capacity is None else capacity
        off = np.zeros(self.n + 1, np.uint32)
        ids = np.zeros(max(cap, 1), np.uint32)
        _chk(lib().u3d_batch_read_packed(self.h, _p(off, _u32), _p(ids, _u32), cap, stream))
        return off, ids[:off[-1]]

    def cull_views(self, packed, stage=STAGE_VISIBLE, stream=None, counts=None, offsets=None, ids=None):
        """One C-ABI call with host buffers in and out (the e2e path)."""
        packed = np.ascontiguousarray(packed)
        self.n = packed.shape[0]
        counts = np.zeros((self.n, 2), np.uint32) if counts is None else counts
        offsets = np.zeros(self.n + 1, np.uint32) if offsets is None else offsets
        ids = np.zeros(max(self.n * self.out_capacity, 1), np.uint32) if ids is None else ids
        _chk(lib().u3d_batch_cull_views(self.h, packed.ctypes.data, self.n, stage, _p(counts, _u32), _p(offsets, _u32), _p(ids, _u32), ids.shape[0],
                                        stream))
        return counts, offsets, ids

    def z_ranges(self, stream=None):
        """(minZ, maxZ) per view after a STAGE_INVIEW run (View::minZ_/maxZ_)."""
        out = np.zeros((self.n, 2), np.float32)
        _chk(lib().u3d_batch_read_z_ranges(self.h, _p(out, _f), stream))
        return out

    def view_ids(self, view, stream=None):
        out = np.zeros(max(self.out_capacity, 1), np.uint32)
        c = C.c_uint32()
        _chk(lib().u3d_batch_read_view(self.h, view, _p(out, _u32), self.out_capacity, C.byref(c), stream))
        return out[:min(c.value, self.out_capacity)].astype(np.int64)

    def depth(self, view, stream=None):
        out = np.zeros((self.height, self.width), np.int32)
        _chk(lib().u3d_batch_read_depth(self.h, view, _p(out, _i32), stream))
        return out

    def mips(self, view, stream=None):
        res = []
        w, h = self.width, self.height
        while True:
            w, h = (w + 1) // 2, (h + 1) // 2
            out = np.zeros((h, w, 2), np.int32)
            _chk(lib().u3d_batch_read_mip(self.h, view, len(res), _p(out, _i32), stream))
            res.append(out)
            if w <= 8 and h <= 8:
                break
        return res

    def tri_stats(self, stream=None):
        out = np.zeros((self.n, 3), np.uint32)
        _chk(lib().u3d_batch_read_tri_stats(self.h, _p(out, _u32), stream))
        return out

    def setup_records(self, view, stream=None):
        c = C.c_uint32()
        _chk(lib().u3d_batch_read_setup_records(self.h, view, None, 0, C.byref(c), stream))
        out = np.zeros((max(c.value, 1), 16), np.int32)
        _chk(lib().u3d_batch_read_setup_records(self.h, view, _p(out, _i32), c.value, C.byref(c), stream))
        return out[:c.value]

    def device_results(self):
        """Raw device pointers (counts [max_views][2], ids [max_views][out_capacity]) for zero-copy use by torch / NCCL."""
        c, i, cap = _vp(), _vp(), C.c_uint32()
        _chk(lib().u3d_batch_device_results(self.h, C.byref(c), C.byref(i), C.byref(cap)))
        return c.value, i.value, cap.value

    def pack_device(self, stream=None):
        """Pack the visible ids on the device; returns device pointers (offsets [n+1], packed ids)."""
        o, p = _vp(), _vp()
        _chk(lib().u3d_batch_pack_device(self.h, stream, C.byref(o), C.byref(p)))
        return o.value, p.value

    def set_occluder_policy(self, size_threshold, max_triangles, camera3, stream=None, enable=True):
        """View::UpdateOccluders heuristics + DrawOccluders' triangle budget.  camera3: n x (halfViewSize, orthoSize, farClip)."""
        cam = None if camera3 is None else np.ascontiguousarray(camera3, np.float32).reshape(-1, 3)
        _chk(lib().u3d_batch_set_occluder_policy(self.h, int(enable), float(size_threshold), int(max_triangles), _p(cam, _f), stream))

    def active_occluders(self, stream=None):
        out = np.zeros(max(self.n, 1), np.uint32)
        _chk(lib().u3d_batch_read_active_occluders(self.h, _p(out, _u32), stream))
        return out[:self.n]

    def occluder_selection(self, view, stream=None):
        cap = 65536
        out = np.zeros(cap, np.uint32)
        c = C.c_uint32()
        _chk(lib().u3d_batch_read_occluder_selection(self.h, view, _p(out, _u32), cap, C.byref(c), stream))
        return out[:c.value].astype(np.int64)

    def process_shadow_casters(self, splits, main_view, main_pos, main_ortho=False, in_view=None, stream=None):
        """View::ProcessShadowCasters over the lists of the last run, in place.  splits: SHADOW_SPLIT_DTYPE array, one per view."""
        sp = np.ascontiguousarray(splits, SHADOW_SPLIT_DTYPE)
        assert sp.shape[0] == self.n
        mv = np.ascontiguousarray(main_view, np.float32).reshape(-1)
        mp = np.ascontiguousarray(main_pos, np.float32)
        iv = None if in_view is None else np.ascontiguousarray(in_view, np.uint8)
        _chk(lib().u3d_batch_process_shadow_casters(self.h, sp.ctypes.data, _p(mv, _f), _p(mp, _f), int(main_ortho),
                                                    None if iv is None else iv.ctypes.data_as(C.POINTER(C.c_uint8)), 0 if iv is None else len(iv), stream))

    def last_launches(self):
        return lib().u3d_batch_last_launches(self.h)

    def set_pipelined(self, on=True):
        """Several batches in flight on different streams: prefer aggregate throughput over single-batch latency (results unchanged)."""
        _chk(lib().u3d_batch_set_pipelined(self.h, 1 if on else 0))

    def close(self):
        if self.h:
            lib().u3d_batch_destroy(self.h)
            self.h = None


class Gather:
    """Result exchange of one batch slot over NVLink peer memory (u3d_gather_*): every rank's packed visible sets and counts land in every
    rank's receive block, written by the producers' own kernel."""

    IPC_HANDLE_BYTES = 64

    def __init__(self, ctx, world, rank, views_per_rank, ids_capacity_per_rank):
        h = _vp()
        _chk(lib().u3d_gather_create(ctx.h, world, rank, views_per_rank, ids_capacity_per_rank, C.byref(h)))
        self.h = h
        self.ctx = ctx
        self.world, self.rank = world, rank
        self.per, self.cap = max(int(views_per_rank), 1), max(int(ids_capacity_per_rank), 1)

    def export(self):
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        _chk(lib().u3d_gather_export(self.h, C.cast(buf, _vp)))
        return buf.raw

    def connect(self, handles):
        """handles: the export() bytes of every rank, in rank order (inter-process, CUDA IPC)."""
        blob = b"".join(handles)
        assert len(blob) == self.world * self.IPC_HANDLE_BYTES
        buf = C.create_string_buffer(blob, len(blob))
        _chk(lib().u3d_gather_connect(self.h, C.cast(buf, _vp)))

    def connect_local(self, gathers):
        """gathers: the Gather objects of every rank, in rank order, all living in this process."""
        arr = (_vp * self.world)(*[g.h for g in gathers])
        _chk(lib().u3d_gather_connect_local(self.h, arr))

    def push(self, batch, stream=None):
        _chk(lib().u3d_batch_push_results(batch.h, self.h, stream))

    def wait(self, stream=None):
        _chk(lib().u3d_gather_wait(self.h, stream))

    def release(self, stream=None):
        _chk(lib().u3d_gather_release(self.h, stream))

    def status(self, stream=None):
        _chk(lib().u3d_gather_status(self.h, stream))

    def device_buffers(self):
        """Device pointers (counts [world][per][2], offsets [world][per + 1], ids [world][cap])."""
        c, o, i = _vp(), _vp(), _vp()
        _chk(lib().u3d_gather_buffers(self.h, C.byref(c), C.byref(o), C.byref(i), None, None))
        return c.value, o.value, i.value

    def read(self, stream=None, ids=True):
        """Host copies of the receive block (synchronises): counts [world, per, 2], offsets [world, per + 1], ids [world, cap]."""
        c = np.zeros((self.world, self.per, 2), np.uint32)
        o = np.zeros((self.world, self.per + 1), np.uint32)
        i = np.zeros((self.world, self.cap), np.uint32) if ids else None
        _chk(lib().u3d_gather_read(self.h, _p(c, _u32), _p(o, _u32), _p(i, _u32), stream))
        return c, o, i

    def close(self):
        if self.h:
            lib().u3d_gather_destroy(self.h)
            self.h = None
