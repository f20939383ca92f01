#!/usr/bin/env python
"""CPU-side statistics of the IsVisible walks on the bench scene (C3): rect extents of the query survivors, how the top-level texels
decide, and how much work three exact strategies need (current DFS refinement, top texels + level-0 scan, plain pixel scan).
usage: python scripts/walk_stats.py [views]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from urho3d_b200 import scenes, capi
from oracle import oraclebind

nviews = int(sys.argv[1]) if len(sys.argv) > 1 else 6
W = H = 256
extent = scenes.CONFIGS["C3"]["extent"]
sc = scenes.Scene(1_000_000, 4000, extent)
views = scenes.agent_cameras(64, extent, W, H)[:nviews]
t = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
t.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
flat = t.flatten()
tree = oraclebind.OracleTree(flat, sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
ob = oraclebind.OracleOB()
ob.set_size(W, H)

def project(vp, boxes):
    """float32 replica of ob_project (good enough for statistics)."""
    n = boxes.shape[0]
    out = np.zeros((n, 6), np.int64)  # early, left, top, right, bottom, z
    mn, mx = boxes[:, :3], boxes[:, 3:]
    xs = []
    early = np.zeros(n, bool)
    minX = maxX = minY = maxY = minZ = None
    for i in range(8):
        p = np.stack([(mx if i & 1 else mn)[:, 0], (mx if i & 2 else mn)[:, 1], (mx if i & 4 else mn)[:, 2]], 1).astype(np.float32)
        c = (p @ vp[:, :3].T + vp[:, 3]).astype(np.float32)
        z = c[:, 2] - np.float32(0.00001)
        early |= z <= 0
        with np.errstate(all="ignore"):
            iw = np.float32(1.0) / c[:, 3]
            sx = iw * c[:, 0] * np.float32(0.5 * W) + np.float32(0.5 * W + 0.5)
            sy = iw * c[:, 1] * np.float32(-0.5 * H) + np.float32(0.5 * H + 0.5)
            sz = iw * z * np.float32(16777216.0)
        if i == 0:
            minX, maxX, minY, maxY, minZ = sx.copy(), sx.copy(), sy.copy(), sy.copy(), sz.copy()
        else:
            minX = np.minimum(minX, sx); maxX = np.maximum(maxX, sx); minY = np.minimum(minY, sy); maxY = np.maximum(maxY, sy); minZ = np.minimum(minZ, sz)
    with np.errstate(all="ignore"):
        left = np.trunc(minX - 1.5); top = np.trunc(minY - 1.5); right = np.round(maxX); bottom = np.round(maxY)
    early |= (right < 0) | (bottom < 0) | (left >= W) | (top >= H) | ~np.isfinite(left + top + right + bottom + minZ)
    out[:, 0] = early
    ok = ~early
    out[ok, 1] = np.clip(left[ok], 0, None); out[ok, 2] = np.clip(top[ok], 0, None)
    out[ok, 3] = np.clip(right[ok], None, W - 1); out[ok, 4] = np.clip(bottom[ok], None, H - 1)
    out[ok, 5] = np.round(minZ[ok]).astype(np.int64) - 16
    return out

tot = dict(boxes=0, early=0, steps_dfs=0, steps_hybrid=0, px_scan=0, top_decided=0, top_vis=0, top_occ=0)
ext_hist = np.zeros(10, np.int64)
for v in views:
    ob.draw_scene_view(sc, v)
    cand = tree.query(1, v.planes, ob, as_ids=False)
    boxes = sc.aabb[tree.order[cand]]
    occl = sc.occludee[tree.order[cand]].astype(bool)
    boxes = boxes[occl]
    vp = (v.proj.reshape(4, 4).astype(np.float32) @ np.vstack([v.view.reshape(3, 4), [0, 0, 0, 1]]).astype(np.float32)).astype(np.float32)
    pr = project(vp, boxes)
    depth = ob.depth().astype(np.int64)
    mips = [m.astype(np.int64) for m in ob.mips()]
    tot["boxes"] += len(pr)
    for e, l, tp, r, b, z in pr:
        if e:
            tot["early"] += 1
            continue
        ext = max(r - l, b - tp) + 1
        ext_hist[min(9, int(ext - 1).bit_length())] += 1
        lv = 0 if ext <= 2 else (int(ext - 1).bit_length() - 1)
        lv = min(lv, len(mips) - 1)
        sh = lv + 1
        # current DFS
        steps = 0
        stack = [(lv, x, y) for y in range(tp >> sh, (b >> sh) + 1) for x in range(l >> sh, (r >> sh) + 1)]
        ntop = len(stack)
        found = False
        top_state = []
        # top-level classification summary
        for (_, x, y) in stack:
            mn_, mx_ = mips[lv][y, x]
            if z <= mn_:
                top_state.append(1)
            elif z > mx_:
                top_state.append(0)
            else:
                s = lv + 1
                px0, py0, px1, py1 = x << s, y << s, min(((x + 1) << s) - 1, W - 1), min(((y + 1) << s) - 1, H - 1)
                top_state.append(1 if (px0 >= l and px1 <= r and py0 >= tp and py1 <= b) else 2)
        if 1 in top_state:
            tot["top_decided"] += 1; tot["top_vis"] += 1
        elif all(s == 0 for s in top_state):
            tot["top_decided"] += 1; tot["top_occ"] += 1
        # DFS as the kernel does it (stack order: last pushed first)
        stack = stack[::-1][::-1]
        st = []
        vis = False
        for ry in range(tp >> sh, (b >> sh) + 1):
            for rx in range(l >> sh, (r >> sh) + 1):
                st = [(lv, rx, ry)]
                while st and not vis:
                    L, x, y = st.pop()
                    steps += 1
                    mn_, mx_ = mips[L][y, x]
                    if z <= mn_:
                        vis = True; break
                    if z > mx_:
                        continue
                    s = L + 1
                    px0, py0, px1, py1 = x << s, y << s, min(((x + 1) << s) - 1, W - 1), min(((y + 1) << s) - 1, H - 1)
                    if px0 >= l and px1 <= r and py0 >= tp and py1 <= b:
                        vis = True; break
                    if L == 0:
                        x0, x1, y0, y1 = max(2 * x, l), min(2 * x + 1, r, W - 1), max(2 * y, tp), min(2 * y + 1, b, H - 1)
                        if (depth[y0:y1 + 1, x0:x1 + 1] >= z).any():
                            vis = True; break
                    else:
                        cl = L - 1
                        x0, x1 = max(2 * x, l >> L), min(2 * x + 1, r >> L, mips[cl].shape[1] - 1)
                        y0, y1 = max(2 * y, tp >> L), min(2 * y + 1, b >> L, mips[cl].shape[0] - 1)
                        for cy in range(y0, y1 + 1):
                            for cx in range(x0, x1 + 1):
                                st.append((cl, cx, cy))
                if vis:
                    break
            if vis:
                break
        tot["steps_dfs"] += steps
        # hybrid: top texels, then level-0 texels of the rect row by row until visible
        hs = ntop
        if not (1 in top_state or all(s == 0 for s in top_state)):
            done = False
            for ty in range(tp >> 1, (b >> 1) + 1):
                for tx in range(l >> 1, (r >> 1) + 1):
                    hs += 1
                    mn_, mx_ = mips[0][ty, tx]
                    if z <= mn_:
                        done = True; break
                    if z > mx_:
                        continue
                    x0, x1, y0, y1 = max(2 * tx, l), min(2 * tx + 1, r), max(2 * ty, tp), min(2 * ty + 1, b)
                    if (depth[y0:y1 + 1, x0:x1 + 1] >= z).any():
                        done = True; break
                if done:
                    break
        tot["steps_hybrid"] += hs
        # plain pixel scan until visible
        sub = depth[tp:b + 1, l:r + 1].reshape(-1) >= z
        tot["px_scan"] += int(np.argmax(sub)) + 1 if sub.any() else sub.size
n = tot["boxes"]
print("views", nviews, "boxes", n, "early-out", tot["early"], "per box: dfs steps %.2f, hybrid steps %.2f, pixel-scan px %.2f" % (
    tot["steps_dfs"] / n, tot["steps_hybrid"] / n, tot["px_scan"] / n))
print("decided by top texels: %.1f%% (visible %.1f%%, occluded %.1f%%)" % (100.0 * tot["top_decided"] / n, 100.0 * tot["top_vis"] / n, 100.0 * tot["top_occ"] / n))
print("rect extent histogram (bit length of extent-1: 0 -> 1px, 1 -> 2, 2 -> 3-4, 3 -> 5-8, 4 -> 9-16, ...):", ext_hist.tolist())
