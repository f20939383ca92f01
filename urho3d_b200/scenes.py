"""Deterministic synthetic scenes for the visibility path (BASELINE.json configs C1..C5).

All arrays are float32 / explicit integer types and are generated from fixed seeds with numpy's MT19937
(np.random.RandomState), so the very same bytes are produced in the build container and on the GPU box.
Nothing here is on the hot path: these are the *inputs* fed identically to the CUDA path, the C oracle
and the compiled reference.
"""
import numpy as np
from . import camera as cam

f32 = np.float32
DRAWABLE_GEOMETRY = 0x1   # GraphicsDefs / Drawable.h:39
CULL_NONE, CULL_CCW, CULL_CW = 0, 1, 2  # GraphicsDefs.h CullMode


def box_mesh(stride=52):
    """A unit box (+-0.5) shaped like bin/Data/Models/Box.mdl as the occlusion path sees it: 24 vertices whose first
    12 bytes are the position, `stride` bytes apart, and 36 16-bit indices.  Faces are wound clockwise
    seen from outside (left-handed, y up), i.e. front faces survive CULL_CCW (OcclusionBuffer.cpp:633-634)."""
    faces = [  # (normal axis, sign)
        (0, +1), (0, -1), (1, +1), (1, -1), (2, +1), (2, -1)]
    verts = []
    idx = []
    for axis, sign in faces:
        n = np.zeros(3)
        n[axis] = sign
        # two tangents so that t1 x t2 = n in a LEFT-handed frame  (=> -n in right-handed maths)
        t1 = np.zeros(3)
        t2 = np.zeros(3)
        t1[(axis + 1) % 3] = 1.0
        t2[(axis + 2) % 3] = 1.0
        if sign > 0:
            t1, t2 = t2, t1
        base = len(verts)
        for a, b in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            verts.append(0.5 * n + 0.5 * a * t1 + 0.5 * b * t2)
        idx += [base, base + 2, base + 1, base, base + 3, base + 2]
    verts = np.asarray(verts, dtype=f32)
    buf = np.zeros((24, stride), dtype=np.uint8)
    buf[:, :12] = verts.view(np.uint8).reshape(24, 12)
    # fill the rest of each vertex with a non-zero pattern so a stride bug cannot go unnoticed
    if stride > 12:
        buf[:, 12:] = 0xA5
    return np.ascontiguousarray(buf), np.asarray(idx, dtype=np.uint16)


def _yaw_models(pos, scale, yaw_deg):
    """3x4 model matrices for translate * rotateY(yaw) * scale, float32."""
    n = pos.shape[0]
    a = (yaw_deg.astype(f32) * f32(np.pi / 180.0)).astype(f32)
    c = np.cos(a).astype(f32)
    s = np.sin(a).astype(f32)
    m = np.zeros((n, 3, 4), dtype=f32)
    m[:, 0, 0] = c * scale[:, 0]
    m[:, 0, 2] = s * scale[:, 2]
    m[:, 1, 1] = scale[:, 1]
    m[:, 2, 0] = -s * scale[:, 0]
    m[:, 2, 2] = c * scale[:, 2]
    m[:, :, 3] = pos
    return m


def _world_aabbs(models, half=0.5):
    """World AABB of a +-half local box under each 3x4 model (centre/extent form, as BoundingBox::Transformed)."""
    centre = models[:, :, 3]
    ext = (np.abs(models[:, :, :3]).sum(axis=2) * f32(half)).astype(f32)
    out = np.zeros((models.shape[0], 6), dtype=f32)
    out[:, :3] = centre - ext
    out[:, 3:] = centre + ext
    return out


class Scene:
    """Drawables (world AABBs + flags) and box occluders."""

    def __init__(self, n_drawables, n_occluders, extent, seed=1234, octree_size=1000.0, octree_levels=8):
        rs = np.random.RandomState(seed)
        E = float(extent)
        self.extent = E
        self.octree_min = np.array([-octree_size] * 3, dtype=f32)
        self.octree_max = np.array([octree_size] * 3, dtype=f32)
        self.octree_levels = octree_levels
        N, K = n_drawables, n_occluders
        scale = rs.uniform(0.5, 2.0, size=(N, 3)).astype(f32)
        pos = np.zeros((N, 3), dtype=f32)
        pos[:, 0] = rs.uniform(-E, E, size=N).astype(f32)
        pos[:, 2] = rs.uniform(-E, E, size=N).astype(f32)
        pos[:, 1] = scale[:, 1] * f32(0.5)
        yaw = rs.uniform(0.0, 360.0, size=N).astype(f32)
        self.aabb = _world_aabbs(_yaw_models(pos, scale, yaw))
        self.flags = np.full(N, DRAWABLE_GEOMETRY, dtype=np.uint8)
        self.view_mask = np.full(N, 0xFFFFFFFF, dtype=np.uint32)
        self.occludee = (rs.uniform(size=N) >= 0.01).astype(np.uint8)
        self.cast_shadows = (rs.uniform(size=N) < 0.5).astype(np.uint8)
        oscale = np.zeros((K, 3), dtype=f32)
        oscale[:, 0] = rs.uniform(8.0, 20.0, size=K).astype(f32)
        oscale[:, 1] = 8.0
        oscale[:, 2] = 1.0
        opos = np.zeros((K, 3), dtype=f32)
        opos[:, 0] = rs.uniform(-E, E, size=K).astype(f32)
        opos[:, 2] = rs.uniform(-E, E, size=K).astype(f32)
        opos[:, 1] = 4.0
        oyaw = rs.uniform(0.0, 360.0, size=K).astype(f32)
        self.occ_model = np.ascontiguousarray(_yaw_models(opos, oscale, oyaw).reshape(K, 12))
        self.occ_aabb = _world_aabbs(self.occ_model.reshape(K, 3, 4))
        self.mesh_vtx, self.mesh_idx = box_mesh(52)
        self.mesh_stride = 52
        self.cull_mode = CULL_CCW

    @property
    def n(self):
        return self.aabb.shape[0]

    @property
    def k(self):
        return self.occ_aabb.shape[0]


def agent_cameras(count, extent, width, height, seed=4321):
    """fov 45, near 0.1, far 2E, position uniform in the region at y=2, yaw uniform, pitch U[-10,10]."""
    rs = np.random.RandomState(seed)
    E = float(extent)
    x = rs.uniform(-E, E, size=count)
    z = rs.uniform(-E, E, size=count)
    yaw = rs.uniform(0.0, 360.0, size=count)
    pitch = rs.uniform(-10.0, 10.0, size=count)
    aspect = float(width) / float(height)
    return [cam.make_view([x[i], 2.0, z[i]], yaw[i], pitch[i], 45.0, aspect, 0.1, 2.0 * E) for i in range(count)]


def shadow_views(extent, n_point_lights=64, seed=4321):
    """C4: 4 orthographic cascades (sizes 50,150,400,1000) + n point lights x 6 cube faces (fov 90, aspect 1)."""
    rs = np.random.RandomState(seed)
    E = float(extent)
    views = []
    for size in (50.0, 150.0, 400.0, 1000.0):
        views.append(cam.make_view([rs.uniform(-E / 2, E / 2), 200.0, rs.uniform(-E / 2, E / 2)], 30.0, 60.0, 45.0, 1.0,
                                   0.0, 1000.0, ortho=True, ortho_size=size))
    face_dirs = [(0.0, 90.0), (0.0, -90.0), (-90.0, 0.0), (90.0, 0.0), (0.0, 0.0), (0.0, 180.0)]  # (pitch, yaw)
    for _ in range(n_point_lights):
        p = [rs.uniform(-E, E), rs.uniform(1.0, 6.0), rs.uniform(-E, E)]
        rng = rs.uniform(20.0, 60.0)
        for pitch, yaw in face_dirs:
            views.append(cam.make_view(p, yaw, pitch, 90.0, 1.0, 0.01 * rng, rng))
    return views


def city_mesh(n_tris=250_000, extent=500.0, seed=1234):
    """C5: one indexed mesh (u32 indices, 12-byte stride) = grid of boxes with random heights, ~n_tris triangles."""
    rs = np.random.RandomState(seed)
    n_boxes = n_tris // 12
    side = int(np.ceil(np.sqrt(n_boxes)))
    bv, bi = box_mesh(12)
    base = bv.view(f32).reshape(24, 3)
    cell = 2.0 * extent / side
    gx, gz = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    gx = gx.reshape(-1)[:n_boxes]
    gz = gz.reshape(-1)[:n_boxes]
    h = rs.uniform(4.0, 40.0, size=n_boxes).astype(f32)
    fp = rs.uniform(0.3, 0.7, size=n_boxes).astype(f32) * f32(cell)
    cx = (-extent + (gx + 0.5) * cell).astype(f32)
    cz = (-extent + (gz + 0.5) * cell).astype(f32)
    v = np.zeros((n_boxes, 24, 3), dtype=f32)
    v[:, :, 0] = base[None, :, 0] * fp[:, None] + cx[:, None]
    v[:, :, 1] = (base[None, :, 1] + f32(0.5)) * h[:, None]
    v[:, :, 2] = base[None, :, 2] * fp[:, None] + cz[:, None]
    idx = (bi.astype(np.uint32)[None, :] + (np.arange(n_boxes, dtype=np.uint32) * 24)[:, None]).reshape(-1)
    return np.ascontiguousarray(v.reshape(-1, 3)), np.ascontiguousarray(idx)


# BASELINE.json configs (SURVEY 8d "Configs as concrete inputs")
CONFIGS = {
    "C1": dict(n=10_000, k=256, extent=150.0, width=256, height=160, views=1),
    "C2": dict(n=1_000_000, k=4000, extent=500.0, width=1024, height=576, views=1),
    "C3": dict(n=1_000_000, k=4000, extent=500.0, width=256, height=256, views=4096),
    "C4": dict(n=2_000_000, k=0, extent=700.0, width=0, height=0, views=388),
    "C5": dict(n=4_000_000, k=0, extent=500.0, width=2048, height=2048, views=1),
}
