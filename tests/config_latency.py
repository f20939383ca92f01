#!/usr/bin/env python
"""Per-phase device time of BASELINE configs C1, C2, C4, C5 (the single-view / frustum-only shapes that are parity cases, not bench
lines), with the compiled reference (oracle/_ref, one host thread) timed beside them.  Diagnostic: prints one JSON line per config.
Lives under tests/ because it executes the checker (oracle/_ref), which only tests/, smoke() and bench.py may do.

    python tests/config_latency.py [C1 C2 C4 C5] [--no-cpu]
"""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from urho3d_b200 import camera, capi, scenes  # noqa: E402


def gpu_phases(batch, reps=7):
    acc = {}
    for _ in range(reps):
        for k, v in batch.run_timed(capi.STAGE_VISIBLE).items():
            acc.setdefault(k, []).append(v)
    med = {k: round(statistics.median(v), 4) for k, v in acc.items()}
    med["total"] = round(sum(med.values()), 4)
    return med


def cpu_view(sc, views, w, h, mesh=None):
    from oracle import refbind
    if not refbind.available():
        return None
    ref = refbind.Ref(0)
    ref.octree_set_size(sc.octree_min, sc.octree_max, sc.octree_levels)
    ref.add_drawables(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    ref.ob_set_size(w, h)
    t0 = time.perf_counter()
    n = 0
    for v in views:
        ref.set_view_raw(v)
        ref.run_view(sc)
        n += 1
        if time.perf_counter() - t0 > 10.0:
            break
    dt = (time.perf_counter() - t0) / n
    ref.close()
    return round(dt * 1e3, 3)


def cpu_shadow_views(sc, views, budget_s=20.0):
    """C4 on the host: the reference's ShadowCasterOctreeQuery (its own text, View.cpp:59-84) per shadow view, one thread, as many of the
    388 views as fit into the time budget (the four cascades first: they dominate)."""
    from oracle import refbind
    if not refbind.available():
        return {}
    ref = refbind.Ref(0)
    ref.octree_set_size(sc.octree_min, sc.octree_max, sc.octree_levels)
    ref.add_drawables(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    t0 = time.perf_counter()
    per = []
    for v in views:
        t1 = time.perf_counter()
        ref.set_view_raw(v)
        ref.query(2)
        per.append(time.perf_counter() - t1)
        if time.perf_counter() - t0 > budget_s:
            break
    ref.close()
    n = len(per)
    # cascades are views 0-3; the cube faces are alike, so the unmeasured ones are taken at the measured faces' mean
    faces = per[4:] or per
    est = sum(per) + (len(views) - n) * (sum(faces) / len(faces))
    return {"cpu_ref_views_timed": n, "cpu_ref_ms_timed": round(sum(per) * 1e3, 2), "cpu_ref_ms_all_views_1thread": round(est * 1e3, 2),
            "cpu_ref_note": "ShadowCasterOctreeQuery per view with the compiled reference, one thread; views beyond the time budget extrapolated from the "
                            "measured cube faces"}


def cpu_city_view(sc, view, w, h, vtx, idx):
    """C5 on the host: the 250k-triangle mesh through the reference's AddTriangles / DrawTriangles / BuildDepthHierarchy, its occluded query and
    the visibility predicate over the query's survivors, one thread."""
    from oracle import refbind
    if not refbind.available():
        return {}
    ref = refbind.Ref(0)
    ref.octree_set_size(sc.octree_min, sc.octree_max, sc.octree_levels)
    ref.add_drawables(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    ref.ob_set_size(w, h)
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    best = None
    for _ in range(2):
        ref.set_view_raw(view)
        t0 = time.perf_counter()
        ref.ob_set_max_triangles(1 << 30)
        ref.ob_clear()
        ref.ob_set_cull_mode(capi.CULL_CCW)
        ref.ob_add_triangles(ident, vtx, 12, idx, 0, idx.shape[0])
        ref.ob_draw()
        t1 = time.perf_counter()
        ref.ob_build_hierarchy()
        t2 = time.perf_counter()
        q = ref.query(1)
        t3 = time.perf_counter()
        vis = ref.check_visibility(True)
        t4 = time.perf_counter()
        cur = {"raster": t1 - t0, "hierarchy": t2 - t1, "query": t3 - t2, "visibility": t4 - t3, "total": t4 - t0}
        if best is None or cur["total"] < best["total"]:
            best = cur
    ref.close()
    return {"cpu_ref_ms_1thread": {k: round(v * 1e3, 3) for k, v in best.items()}, "cpu_ref_query_result": int(len(q)), "cpu_ref_visible": int(len(vis))}


def run(name, cpu):
    cfg = scenes.CONFIGS[name]
    ctx = capi.Context(0)
    for kv in filter(None, os.environ.get("U3D_DEBUG_OPTS", "").split(",")):  # tuning sweeps only, e.g. "tile_target=444"
        opt, val = kv.split("=")
        capi.debug_set_option(opt, int(val))
    out = {"config": name}
    w, h = cfg["width"], cfg["height"]
    if name == "C5":
        vtx, idx = scenes.city_mesh(250_000, cfg["extent"])
        sc = scenes.Scene(cfg["n"], 4, cfg["extent"], seed=55)
        mesh = capi.Mesh(ctx, vtx, 12, idx)
        ident = np.array([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]], np.float32)
        occ = capi.Occluders(ctx, np.array([[-500, 0, -500, 500, 40, 500]], np.float32), ident, [mesh], cull_mode=capi.CULL_CCW)
        views = [camera.make_view([20.0, 30.0, -480.0], 10.0, 12.0, 50.0, 1.0, 0.5, 1200.0)]
        packed = capi.pack_views(views)
    elif name == "C4":
        sc = scenes.Scene(cfg["n"], 4, cfg["extent"], seed=4)
        views = scenes.shadow_views(cfg["extent"], n_point_lights=64)
        mesh = occ = None
        packed = capi.pack_views(views, query_mode=capi.QUERY_SHADOWCASTER, use_occlusion=False)
    else:
        sc = scenes.Scene(cfg["n"], cfg["k"], cfg["extent"])
        views = scenes.agent_cameras(cfg["views"], cfg["extent"], w, h)
        mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
        occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
        packed = capi.pack_views(views)
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    t0 = time.perf_counter()
    gs = capi.GpuScene(ctx, octree=tree)
    out["flatten_upload_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    batch = capi.Batch(ctx, gs, occ, w, h, len(views), sc.n if len(views) < 16 else 1 << 20)
    batch.set_views(packed)
    for _ in range(3):
        batch.run(capi.STAGE_VISIBLE)
    counts = batch.counts()
    out["views"] = len(views)
    out["drawables"] = sc.n
    out["octants"] = gs.num_octants
    out["query_result"] = float(counts[:, 0].mean())
    out["visible"] = float(counts[:, 1].mean())
    if occ is not None:
        out["submitted_tris"] = float(batch.tri_stats()[:, 0].mean())
    out["gpu_ms"] = gpu_phases(batch)
    # the whole call with host buffers (one view: the latency a View::GetDrawables call site would see)
    t = []
    for _ in range(7):
        t0 = time.perf_counter()
        batch.cull_views(packed, capi.STAGE_VISIBLE)
        t.append(time.perf_counter() - t0)
    out["call_ms"] = round(statistics.median(t) * 1e3, 3)
    if cpu and name == "C4":
        out.update(cpu_shadow_views(sc, views))
    elif cpu and name == "C5":
        out.update(cpu_city_view(sc, views[0], w, h, vtx, idx))
    elif cpu:
        out["cpu_ref_ms_per_view_1thread"] = cpu_view(sc, views, w, h)
    print(json.dumps(out), flush=True)
    for o in (batch, occ, mesh, gs, tree, ctx):
        if o is not None:
            o.close()


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or ["C1", "C2", "C4", "C5"]
    for nm in names:
        run(nm, "--no-cpu" not in sys.argv)
