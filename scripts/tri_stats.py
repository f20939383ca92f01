"""Dump set-up triangle size statistics of the C3 workload (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from urho3d_b200 import capi, scenes
sc = scenes.Scene(1000, 4000, 500.0)
views = scenes.agent_cameras(128, 500.0, 256, 256)
ctx = capi.Context(0)
tree = capi.Octree(); tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
gs = capi.GpuScene(ctx, octree=tree)
mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
b = capi.Batch(ctx, gs, occ, 256, 256, len(views), 2000)
b.set_views(capi.pack_views(views)); b.run(); b.counts()
H = []; W = []; PX = []; N = []
for v in range(len(views)):
    r = b.setup_records(v).astype(np.int64)
    N.append(len(r))
    if not len(r): continue
    y0, y1, y2 = r[:, 0], r[:, 1], r[:, 2]
    h = y2 - y0
    # spans at every row (vectorised per half)
    px = np.zeros(len(r), np.int64); wmax = np.zeros(len(r), np.int64)
    for (ya, yb, lx, lxs, rx, rxs) in ((y0, y1, r[:, 4], r[:, 5], r[:, 8], r[:, 9]), (y1, y2, r[:, 10], r[:, 11], r[:, 14], r[:, 15])):
        nrow = np.maximum(yb - ya, 0)
        for k in range(int(nrow.max()) if len(nrow) else 0):
            m = k < nrow
            xl = (lx + k * lxs) >> 16; xr = (rx + k * rxs) >> 16
            s = np.where(m, np.maximum(xr - xl, 0), 0)
            px += s; wmax = np.maximum(wmax, s)
    H.append(h); W.append(wmax); PX.append(px)
H = np.concatenate(H); W = np.concatenate(W); PX = np.concatenate(PX)
print("views", len(views), "tris/view mean", np.mean(N), "max", np.max(N))
print("pixels/view mean", PX.sum() / len(views))
area = H * np.maximum(W, 1)
for name, a in (("rows", H), ("maxspan", W), ("pixels", PX), ("bbox", area)):
    print(name, "pct 50/90/99/max", np.percentile(a, [50, 90, 99]).tolist(), a.max())
bins = [0, 16, 64, 256, 1024, 4096, 16384, 1 << 30]
for lo, hi in zip(bins[:-1], bins[1:]):
    m = (area > lo) & (area <= hi) if lo else (area <= hi)
    print("bbox area (%d,%d]: tris %.1f%%  pixels %.1f%%" % (lo, hi, 100 * m.mean(), 100 * PX[m].sum() / max(PX.sum(), 1)))
