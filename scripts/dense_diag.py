"""Cull part of the C3 batch alone (no pipelining): time per path, and how evenly the dense path's warps finish (run on the GPU box).
usage: python scripts/dense_diag.py [views]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from urho3d_b200 import capi, scenes
V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenes.Scene(1_000_000, 4000, 500.0)
views = scenes.agent_cameras(V, 500.0, 256, 256)
ctx = capi.Context(0)
for kv in filter(None, os.environ.get("U3D_DEBUG_OPTS", "").split(",")):
    name, val = kv.split("=")
    capi.debug_set_option(name, int(val))
tree = capi.Octree(); tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
gs = capi.GpuScene(ctx, octree=tree)
mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
b = capi.Batch(ctx, gs, occ, 256, 256, V, 32768, tri_pool=V * 14000)
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
for name, mode in (("per-view kernel", 0), ("dense path", 2)):
    capi.debug_set_option("cull_dense", mode)
    print("%-16s cull visible %.3f ms   query only %.3f ms" % (name, timeit(lambda: b.run(capi.STAGE_VISIBLE, s, part=2)), timeit(lambda: b.run(capi.STAGE_QUERY, s, part=2))))
capi.debug_set_option("read_dense_prof", 1)
capi.debug_set_option("dense_prof", 1)
b.run(capi.STAGE_VISIBLE, s, part=2); torch.cuda.synchronize()
p = capi.read_cull_prof().astype(np.float64)
for k, name in ((0, "level kernels (all levels together)"), (1, "occlusion kernel")):
    t0, t1, busy, n = p[4 * k:4 * k + 4]
    if n:
        print("%-36s span %.3f ms, warps %d, mean busy %.3f ms = %.1f%% of the span" % (name, (t1 - t0) * 1e-6, n, busy / n * 1e-6, 100.0 * busy / n / (t1 - t0)))
t_first = None
for lv in range(1, 10):
    capi.debug_set_option("read_dense_prof", 2 + lv)  # that level's four numbers
    q = capi.read_cull_prof().astype(np.float64)
    t0, t1, busy, n = q[:4]
    if n:
        t_first = t0 if t_first is None else t_first
        print("  level %d: starts at %.3f ms, span %.3f ms, warps %d, mean busy %.3f ms (%.0f%%)" % (lv, (t0 - t_first) * 1e-6, (t1 - t0) * 1e-6, n, busy / n * 1e-6, 100.0 * busy / n / (t1 - t0)))
capi.debug_set_option("dense_prof", 0)
