"""Time the cull kernel alone in several modes to see where its time goes (run on the GPU box)."""
import sys, os, time
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
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b.set_views(capi.pack_views(views), s); b.run(capi.STAGE_VISIBLE, s); c = b.counts(s)
print("views", V, "query/visible per view", c.mean(axis=0))
print("full pipeline ms", timeit(lambda: b.run(capi.STAGE_VISIBLE, s)))
print("raster part ms", timeit(lambda: b.run(capi.STAGE_VISIBLE, s, part=1)))
print("cull visible (occluded query + predicate) ms", timeit(lambda: b.run(capi.STAGE_VISIBLE, s, part=2)))
print("cull query only (occluded query) ms", timeit(lambda: b.run(capi.STAGE_QUERY, s, part=2)))
capi.debug_set_option("cull_prof", 1)
b.run(capi.STAGE_VISIBLE, s, part=2); torch.cuda.synchronize()
prof = capi.read_cull_prof().astype(np.float64)
capi.debug_set_option("cull_prof", 0)
names = ["prologue", "octant sweeps", "octant tests lv>=7", "chunk set-up", "frustum rounds", "drawable tests+emit", "octant tests lv<=5", "octant tests lv 6"]
print("in-kernel phase share (thread-0 clocks summed over CTAs):")
for n, p in zip(names, prof):
    if p: print("   %-22s %5.1f%%   %8.0f cycles per view" % (n, 100 * p / prof.sum(), p / V))
b.set_views(capi.pack_views(views, query_mode=capi.QUERY_FRUSTUM), s)
print("frustum query + predicate ms", timeit(lambda: b.run(capi.STAGE_VISIBLE, s, part=2)))
print("frustum query only ms", timeit(lambda: b.run(capi.STAGE_QUERY, s, part=2)))
pass
b.set_views(capi.pack_views(views, query_mode=capi.QUERY_FRUSTUM, flags=0), s)
print("frustum query, nothing passes the flag filter ms", timeit(lambda: b.run(capi.STAGE_QUERY, s, part=2)))
