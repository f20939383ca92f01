"""Sweep the size-class thresholds of the tile rasteriser (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from urho3d_b200 import capi, scenes
V = 2048
sc = scenes.Scene(1000, 4000, 500.0)
views = scenes.agent_cameras(V, 500.0, 256, 256)
ctx = capi.Context(0)
tree = capi.Octree(); tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
gs = capi.GpuScene(ctx, octree=tree)
mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
b = capi.Batch(ctx, gs, occ, 256, 256, V, 2000, tri_pool=V * 14000)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
b.set_views(capi.pack_views(views), s)
for lane in (16, 64, 256):
    for warp in (256, 512, 1024, 2048, 8192):
        capi.debug_set_option("lane_area_max", lane); capi.debug_set_option("warp_area_max", warp)
        for _ in range(2): b.run_timed(capi.STAGE_VISIBLE, s)
        ms = [b.run_timed(capi.STAGE_VISIBLE, s) for _ in range(5)]
        print("lane<=%4d warp<=%5d  setup %.3f  fill %.3f ms" % (lane, warp, np.median([m["setup"] for m in ms]), np.median([m["fill"] for m in ms])))
