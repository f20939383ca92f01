#!/usr/bin/env python
"""Cost of the ranked occluder selection (u3d_batch_set_occluder_policy) on the bench workload: per-phase device time of the C3 batch with
the default rule (every occluder in the frustum) and with the reference's heuristics (size threshold 0.025, 5000-triangle budget)."""
import json
import os
import statistics
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from urho3d_b200 import capi, scenes  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = scenes.CONFIGS["C3"]
sc = scenes.Scene(cfg["n"], cfg["k"], cfg["extent"])
views = scenes.agent_cameras(V, cfg["extent"], 256, 256)
ctx = capi.Context(0)
tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
gs = capi.GpuScene(ctx, octree=tree)
mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
batch = capi.Batch(ctx, gs, occ, 256, 256, V, 32768, tri_pool=V * 14000)
batch.set_views(capi.pack_views(views))
cam3 = np.array([[v.half_view_size, v.ortho_size, v.far] for v in views], np.float32)
for name, policy in (("default rule", None), ("heuristics thr=0.025 budget=5000, threaded branch", (0.025, 5000, 1)),
                     ("heuristics thr=0.025 budget=5000, non-threaded branch (serial per view)", (0.025, 5000, 2))):
    if policy:
        batch.set_occluder_policy(policy[0], policy[1], cam3, enable=policy[2])
    acc = {}
    for _ in range(5):
        for k, ms in batch.run_timed(capi.STAGE_VISIBLE).items():
            acc.setdefault(k, []).append(ms)
    counts = batch.counts()
    stats = batch.tri_stats()
    print(json.dumps({"selection": name, "views": V, "phases_ms": {k: round(statistics.median(v), 4) for k, v in acc.items()},
                      "total_ms": round(sum(statistics.median(v) for v in acc.values()), 3), "submitted_tris_per_view": float(stats[:, 0].mean()),
                      "query_result_per_view": float(counts[:, 0].mean()), "visible_per_view": float(counts[:, 1].mean())}))
