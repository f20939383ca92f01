"""Sweep the rect size from which a box is walked by a whole warp in the cull kernel (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from urho3d_b200 import capi, scenes
V = 2048
sc = scenes.Scene(1_000_000, 4000, 500.0)
views = scenes.agent_cameras(V, 500.0, 256, 256)
ctx = capi.Context(0)
tree = capi.Octree(); tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
gs = capi.GpuScene(ctx, octree=tree)
mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
b = capi.Batch(ctx, gs, occ, 256, 256, V, 200000, tri_pool=V * 14000)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
b.set_views(capi.pack_views(views), s); b.run(capi.STAGE_VISIBLE, s); ref = b.counts(s).copy()
for px in (16, 32, 48, 64, 96, 128, 192, 300):
    capi.debug_set_option("heavy_extent", px)
    for _ in range(2): b.run(capi.STAGE_VISIBLE, s, part=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): b.run(capi.STAGE_VISIBLE, s, part=2)
    e1.record(); torch.cuda.synchronize()
    assert np.array_equal(b.counts(s), ref)
    print("heavy extent >= %3d px: cull %.3f ms" % (px, e0.elapsed_time(e1) / 5))
