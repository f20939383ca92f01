"""GPU parity tests: every call goes through the C-ABI of libu3dgpu.so and is compared bit-for-bit with the C oracle
(oracle/u3d_oracle.c, itself pinned to the compiled reference) and with the golden fixtures."""
import glob
import os

import numpy as np
import pytest

import helpers
from oracle import oraclebind
from urho3d_b200 import camera, capi, scenes

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("cull_path")]

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def gpu_scene(ctx, sc):
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    flat = tree.flatten()
    return tree, flat, capi.GpuScene(ctx, octree=tree)


def oracle_view(sc, otree, ob, v):
    sub = ob.draw_scene_view(sc, v)
    cand = otree.query(1, v.planes, ob, as_ids=False)
    vis = otree.order[otree.check_visibility(ob, cand)]
    return sub, otree.order[cand], vis


@pytest.mark.parametrize("w,h,stress", [(256, 160, False), (128, 128, True), (64, 30, True), (512, 288, False), (32, 32, True)])
def test_occlusion_buffer_dropin(gpu_ctx, w, h, stress):
    """OcclusionBuffer API call by call: SetSize/SetView/Clear/SetCullMode/AddTriangles/DrawTriangles/BuildDepthHierarchy/IsVisible."""
    sc = helpers.small_scene(n=4000, k=128, extent=90.0, seed=3 + w)
    views = helpers.stress_views(90.0, w, h, 6, seed=w + h) if stress else scenes.agent_cameras(6, 90.0, w, h, seed=w)
    gob = capi.OcclusionBuffer(gpu_ctx)
    assert gob.IsVisible(sc.aabb[0])  # no buffer yet -> true (OB.cpp:338-339)
    assert gob.SetSize(w, h)
    ob = oraclebind.OracleOB()
    ob.set_size(w, h)
    assert (gob.GetWidth(), gob.GetHeight()) == ob.size()
    for v in views:
        gob.SetView(v)
        ob.set_view(v)
        assert np.array_equal(gob.GetViewProj(), ob.viewproj())
        gob.SetMaxTriangles(1 << 30)
        ob.set_max_triangles(1 << 30)
        gob.Clear()
        ob.clear()
        for i in range(sc.k):
            if oraclebind.frustum_test(v.planes, sc.occ_aabb[i], fast=True) == 0:
                continue
            gob.SetCullMode(sc.cull_mode)
            ob.set_cull_mode(sc.cull_mode)
            a = gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx, 0, 36)
            b = ob.draw_batch(sc.occ_model[i], sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx, 0, 36)
            assert a == b
        gob.DrawTriangles()
        assert np.array_equal(gob.GetBuffer(), ob.depth())
        assert gob.GetNumTriangles() == ob.num_triangles()
        # hierarchy still dirty: pixel-scan path
        assert np.array_equal(gob.IsVisibleMany(sc.aabb[:1500]), ob.is_visible_many(sc.aabb[:1500]))
        gob.BuildDepthHierarchy()
        ob.build_hierarchy()
        for a, b in zip(gob.GetMips(), ob.mips()):
            assert np.array_equal(a, b)
        assert np.array_equal(gob.IsVisibleMany(sc.aabb), ob.is_visible_many(sc.aabb))
        assert gob.IsVisible(sc.aabb[7]) == ob.is_visible(sc.aabb[7])
    gob.close()
    ob.close()


def test_set_size_rules(gpu_ctx):
    gob = capi.OcclusionBuffer(gpu_ctx)
    assert not gob.SetSize(100, 64)      # not a power of two (OB.cpp:73-77)
    assert not gob.SetSize(0, 64)
    assert not gob.SetSize(64, -2)
    assert gob.SetSize(64, 33)
    assert gob.GetHeight() == 34         # odd heights are rounded up (OB.cpp:63-65)
    assert gob.SetSize(64, 34)
    gob.close()


def test_budget_cullmodes_nonindexed_u32(gpu_ctx):
    w, h = 64, 64
    sc = helpers.small_scene(n=10, k=8, extent=10.0, seed=9)
    v = helpers.stress_views(10.0, w, h, 1, seed=4)[0]
    pos = sc.mesh_vtx[:, :12].copy().view(np.float32).reshape(24, 3)
    flat_vtx = np.ascontiguousarray(pos[sc.mesh_idx.astype(np.int64)])
    idx32 = sc.mesh_idx.astype(np.uint32)
    gob = capi.OcclusionBuffer(gpu_ctx)
    gob.SetSize(w, h)
    ob = oraclebind.OracleOB()
    ob.set_size(w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, 52, sc.mesh_idx)
    for cull in (0, 1, 2):
        for reverse in (False, True):
            v.reverse_culling = reverse
            gob.SetView(v)
            ob.set_view(v)
            gob.SetMaxTriangles(20)
            ob.set_max_triangles(20)
            gob.Clear()
            ob.clear()
            for i in range(sc.k):
                gob.SetCullMode(cull)
                ob.set_cull_mode(cull)
                if i % 4 == 0:
                    a = gob.AddTriangles(sc.occ_model[i], flat_vtx, 12, None, 3, 33)
                    b = ob.draw_batch(sc.occ_model[i], flat_vtx, 12, None, 3, 33)
                elif i % 4 == 1:
                    a = gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, 52, idx32, 6, 30)
                    b = ob.draw_batch(sc.occ_model[i], sc.mesh_vtx, 52, idx32, 6, 30)
                elif i % 4 == 2:
                    a = gob.AddTrianglesMesh(sc.occ_model[i], mesh, 3, 33)
                    b = ob.draw_batch(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 3, 33)
                else:
                    a = gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36)
                    b = ob.draw_batch(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36)
                assert a == b
            assert gob.GetCullMode() == ob.cull_mode()
            gob.DrawTriangles()
            # a second DrawTriangles with nothing queued must not change anything (batches_ cleared, OB.cpp:231)
            gob.DrawTriangles()
            assert np.array_equal(gob.GetBuffer(), ob.depth())
            assert gob.GetNumTriangles() == ob.num_triangles()
    mesh.close()
    gob.close()
    ob.close()


def test_reset(gpu_ctx):
    """OcclusionBuffer::Reset (OB.cpp:146-150) through u3d_ob_reset: queued batches are dropped undrawn, what was drawn stays, numTriangles_
    and with it the budget start again (CPU pin of the same sequence: test_reset_drops_queued_batches_and_the_triangle_count)."""
    w, h = 64, 64
    sc = helpers.small_scene(n=10, k=8, extent=10.0, seed=9)
    v = helpers.stress_views(10.0, w, h, 1, seed=4)[0]
    gob = capi.OcclusionBuffer(gpu_ctx)
    gob.SetSize(w, h)
    ob = oraclebind.OracleOB()
    ob.set_size(w, h)
    gob.SetView(v)
    ob.set_view(v)
    gob.SetMaxTriangles(30)
    ob.set_max_triangles(30)
    gob.Clear()
    ob.clear()
    gob.SetCullMode(sc.cull_mode)
    ob.set_cull_mode(sc.cull_mode)
    assert gob.AddTriangles(sc.occ_model[0], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36) == ob.draw_batch(sc.occ_model[0], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36)
    gob.DrawTriangles()
    gob.Reset()
    ob.reset()
    assert gob.GetNumTriangles() == 0 == ob.num_triangles()
    assert np.array_equal(gob.GetBuffer(), ob.depth())
    for i in (1, 2, 3):
        gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36)
    gob.Reset()
    gob.DrawTriangles()
    assert gob.GetNumTriangles() == 0
    assert np.array_equal(gob.GetBuffer(), ob.depth())
    rets = []
    for i in (4, 5, 6, 7):
        rets.append((gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36), ob.draw_batch(sc.occ_model[i], sc.mesh_vtx, 52, sc.mesh_idx, 0, 36)))
    gob.DrawTriangles()
    assert all(a == b for a, b in rets) and rets[0][0] and not rets[-1][0]
    assert np.array_equal(gob.GetBuffer(), ob.depth())
    assert gob.GetNumTriangles() == ob.num_triangles()
    gob.BuildDepthHierarchy()
    ob.build_hierarchy()
    for a, b in zip(gob.GetMips(), ob.mips()):
        assert np.array_equal(a, b)
    gob.close()
    ob.close()


@pytest.mark.parametrize("mode", [capi.QUERY_FRUSTUM, capi.QUERY_OCCLUDED, capi.QUERY_SHADOWCASTER])
def test_octree_queries(gpu_ctx, mode):
    """Octree::GetDrawables(query) result_ sequence for the three frustum query classes, with flag/mask filtering."""
    w, h = 256, 160
    sc = helpers.small_scene(n=6000, k=128, extent=100.0, seed=31)
    sc.flags[::7] = 0x2          # DRAWABLE_LIGHT: filtered by drawableFlags
    sc.view_mask[::5] = 0x4
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    gob = capi.OcclusionBuffer(gpu_ctx)
    gob.SetSize(w, h)
    for v in scenes.agent_cameras(5, 100.0, w, h, seed=8):
        ob.draw_scene_view(sc, v)
        gob.SetView(v)
        gob.SetMaxTriangles(1 << 30)
        gob.Clear()
        gob.SetCullMode(sc.cull_mode)
        for i in range(sc.k):
            if oraclebind.frustum_test(v.planes, sc.occ_aabb[i], fast=True):
                gob.AddTriangles(sc.occ_model[i], sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx, 0, 36)
        gob.DrawTriangles()
        gob.BuildDepthHierarchy()
        for flags, mask in ((0xFF, 0xFFFFFFFF), (0x1, 0xFFFFFFFF), (0xFF, 0x3), (0x2, 0x4)):
            want = otree.query(mode, v.planes, ob if mode == 1 else None, flags, mask)
            got, qc = gs.query(v.planes, gob if mode == 1 else None, flags, mask, mode, capi.STAGE_QUERY)
            assert qc == len(want)
            assert np.array_equal(got, want)
            if mode == 1:
                cand = otree.query(1, v.planes, ob, flags, mask, as_ids=False)
                vis = otree.order[otree.check_visibility(ob, cand)]
                got, qc = gs.query(v.planes, gob, flags, mask, mode, capi.STAGE_VISIBLE)
                assert qc == len(cand) and np.array_equal(got, vis)
    with pytest.raises(capi.U3DError) as e:
        gs.query(v.planes, None, mode=capi.QUERY_FRUSTUM, capacity=1)
    assert e.value.code == capi.ERR_CAPACITY
    for o in (gob, gs, tree):
        o.close()
    ob.close()


def test_cull_views_with_a_blocking_wait(gpu_ctx):
    """`blocking_wait` (for callers with more threads than host cores) changes how the call waits, not what it returns."""
    w = h = 128
    sc = helpers.small_scene(n=6000, k=120, extent=100.0, seed=77)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = scenes.agent_cameras(24, 100.0, w, h, seed=5)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    packed = capi.pack_views(views)
    ref = [np.array(a, copy=True) for a in batch.cull_views(packed, capi.STAGE_VISIBLE)]
    capi.debug_set_option("blocking_wait", 1)
    try:
        for _ in range(3):
            got = batch.cull_views(packed, capi.STAGE_VISIBLE)
            for a, b in zip(ref, got):
                assert np.array_equal(a, b)
    finally:
        capi.debug_set_option("blocking_wait", 0)
    assert ref[0][:, 1].sum() > 0
    for o in (batch, occ, mesh, gs, tree):
        o.close()


@pytest.mark.parametrize("w,h,nviews,stress", [(256, 160, 12, False), (256, 256, 8, True), (64, 64, 40, True), (1024, 576, 2, False)])
def test_batched_views(gpu_ctx, w, h, nviews, stress):
    """The batched many-view entry point: depth plane, every mip level, tri counts, query result and visible set per view."""
    sc = helpers.small_scene(n=8000, k=160, extent=100.0, seed=w + 1)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = helpers.stress_views(100.0, w, h, nviews, seed=h) if stress else scenes.agent_cameras(nviews, 100.0, w, h, seed=h)
    views[1].reverse_culling = True
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, nviews, sc.n)
    counts, offsets, ids = batch.cull_views(capi.pack_views(views), capi.STAGE_VISIBLE)
    stats = batch.tri_stats()
    for i, v in enumerate(views):
        sub, cand, vis = oracle_view(sc, otree, ob, v)
        assert np.array_equal(batch.depth(i), ob.depth()), "view %d depth" % i
        for a, b in zip(batch.mips(i), ob.mips()):
            assert np.array_equal(a, b)
        assert stats[i, 0] == sub and stats[i, 1] == ob.num_triangles()
        assert counts[i, 0] == len(cand) and counts[i, 1] == len(vis)
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], vis)
        assert np.array_equal(batch.view_ids(i), vis)
    # query stage only
    batch.set_views(capi.pack_views(views))
    batch.run(capi.STAGE_QUERY)
    off, qids = batch.packed()
    for i, v in enumerate(views):
        sub, cand, vis = oracle_view(sc, otree, ob, v)
        assert np.array_equal(qids[off[i]:off[i + 1]], cand)
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


@pytest.mark.parametrize("opt", ["dense_queue_cap", "dense_pool_cap"])
def test_dense_path_overflow_is_answered_on_the_device(gpu_ctx, opt):
    """The batch-wide cull path works out of bounded scratch (a queue of (view, octant) pairs, a pool of result words).  When either is
    too small the sequence raises a device flag and the one-CTA-per-view kernel queued behind it produces the result: same sequences,
    no error, no host round trip."""
    w, h, nviews = 128, 128, 24
    sc = helpers.small_scene(n=8000, k=96, extent=100.0, seed=77)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = scenes.agent_cameras(nviews, 100.0, w, h, seed=3)
    capi.debug_set_option("cull_dense", 2)
    capi.debug_set_option(opt, 64)
    try:
        batch = capi.Batch(gpu_ctx, gs, occ, w, h, nviews, sc.n)
        for stage in (capi.STAGE_VISIBLE, capi.STAGE_QUERY):
            counts, offsets, ids = batch.cull_views(capi.pack_views(views), stage)
            for i, v in enumerate(views):
                sub, cand, vis = oracle_view(sc, otree, ob, v)
                want = vis if stage == capi.STAGE_VISIBLE else cand
                assert counts[i, 0] == len(cand) and np.array_equal(ids[offsets[i]:offsets[i + 1]], want)
        batch.close()
    finally:
        capi.debug_set_option(opt, 0)
    for o in (occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_frustum_only_shadow_views(gpu_ctx):
    """BASELINE config 4 shape: ortho cascades + cube faces, no occlusion buffer, ShadowCasterOctreeQuery."""
    sc = helpers.small_scene(n=20000, k=4, extent=120.0, seed=77)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, _ = helpers.make_oracle(sc, flat, 8, 8)
    views = scenes.shadow_views(120.0, n_point_lights=6, seed=3)
    batch = capi.Batch(gpu_ctx, gs, None, 0, 0, len(views), sc.n)
    counts, offsets, ids = batch.cull_views(capi.pack_views(views, query_mode=capi.QUERY_SHADOWCASTER, use_occlusion=False), capi.STAGE_VISIBLE)
    total = 0
    for i, v in enumerate(views):
        want = otree.query(2, v.planes, None)
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], want)
        total += len(want)
    assert total > 0
    for o in (batch, gs, tree):
        o.close()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures_through_gpu(gpu_ctx, path):
    """Committed outputs of the UNMODIFIED reference vs the CUDA path."""
    g = np.load(path)
    sc = scenes.Scene(int(g["n"]), int(g["k"]), float(g["extent"]), seed=int(g["seed"]))
    w, h = int(g["width"]), int(g["height"])
    views = []
    for i in range(int(g["nviews"])):
        nf = g["v%d_nearfar" % i]
        views.append(camera.CameraView(g["v%d_view" % i], g["v%d_proj" % i], g["v%d_planes" % i], nf[0], nf[1]))
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    for key in ("parent", "level", "first", "count", "order", "box"):
        assert np.array_equal(flat[key], g["flat_" + key])
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    counts, offsets, ids = batch.cull_views(capi.pack_views(views), capi.STAGE_VISIBLE)
    stats = batch.tri_stats()
    for i in range(len(views)):
        assert np.array_equal(batch.depth(i), g["v%d_depth" % i])
        for lv, m in enumerate(batch.mips(i)):
            assert np.array_equal(m, g["v%d_mip%d" % (i, lv)])
        assert stats[i, 1] == int(g["v%d_numtris" % i][0])
        assert counts[i, 0] == len(g["v%d_query" % i])
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], g["v%d_visible" % i])
    for o in (batch, occ, mesh, gs, tree):
        o.close()


def test_million_drawables_slice(gpu_ctx):
    """BASELINE configs 2/3 scene (1M drawables, 4000 occluders): one 1024x576 view and four 256x256 views against the oracle."""
    cfg = scenes.CONFIGS["C3"]
    sc = scenes.Scene(cfg["n"], cfg["k"], cfg["extent"])
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    assert gs.num_drawables == cfg["n"]
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    for (w, h, nv) in ((256, 256, 4), (1024, 576, 1)):
        otree, ob = helpers.make_oracle(sc, flat, w, h)
        views = scenes.agent_cameras(nv, cfg["extent"], w, h)
        batch = capi.Batch(gpu_ctx, gs, occ, w, h, nv, 200000)
        counts, offsets, ids = batch.cull_views(capi.pack_views(views), capi.STAGE_VISIBLE)
        for i, v in enumerate(views):
            sub, cand, vis = oracle_view(sc, otree, ob, v)
            assert np.array_equal(batch.depth(i), ob.depth())
            for a, b in zip(batch.mips(i), ob.mips()):
                assert np.array_equal(a, b)
            assert counts[i, 0] == len(cand)
            assert np.array_equal(ids[offsets[i]:offsets[i + 1]], vis)
        batch.close()
        ob.close()
    for o in (occ, mesh, gs, tree):
        o.close()


def test_general_path_for_every_triangle(gpu_ctx):
    """Force every triangle through the irregular (global-memory, linear-addressing) path of the batched pipeline."""
    w, h = 256, 160
    sc = helpers.small_scene(n=3000, k=128, extent=80.0, seed=5)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = helpers.stress_views(80.0, w, h, 6, seed=12)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    capi.debug_set_option("force_irregular", 1)
    try:
        counts, offsets, ids = batch.cull_views(capi.pack_views(views), capi.STAGE_VISIBLE)
    finally:
        capi.debug_set_option("force_irregular", 0)
    for i, v in enumerate(views):
        sub, cand, vis = oracle_view(sc, otree, ob, v)
        assert np.array_equal(batch.depth(i), ob.depth())
        for a, b in zip(batch.mips(i), ob.mips()):
            assert np.array_equal(a, b)
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], vis)
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_garbage_geometry_does_not_crash_and_matches(gpu_ctx):
    """Huge / non-finite / degenerate vertices: conversions follow x86 (INT_MIN), spans may spill across rows (linear addressing),
    nothing may hang or fault, and whatever lands inside the plane must equal the oracle."""
    w, h = 64, 64
    rs = np.random.RandomState(3)
    v = camera.make_view([0, 0, 0], 0, 0, 60.0, 1.0, 0.1, 100.0)
    base = rs.uniform(-3, 3, size=(60, 3)).astype(np.float32)
    base[:, 2] += 6.0
    vtx = base.copy()
    vtx[3] = [1e30, 0, 5]
    vtx[7] = [0, -1e30, 5]
    vtx[11] = [np.inf, 1, 5]
    vtx[13] = [np.nan, 1, 5]
    vtx[17] = [1e20, 1e20, 1e20]
    vtx[20] = vtx[21]                      # zero-area
    vtx[25:28] = [[0, 0, 1e-30]] * 3       # at the eye
    vtx[30] = [1e6, -1e6, 0.2]             # sliver through the near plane
    model = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], np.float32)
    tree = capi.Octree()
    tree.add(np.array([[-1, -1, 5, 1, 1, 7]], np.float32))
    gs = capi.GpuScene(gpu_ctx, octree=tree)
    mesh = capi.Mesh(gpu_ctx, vtx, 12, None)
    occ = capi.Occluders(gpu_ctx, np.array([[-1, -1, 5, 1, 1, 7]], np.float32), model.reshape(1, 12), [mesh], cull_mode=0)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, 1, 16)
    batch.cull_views(capi.pack_views([v]), capi.STAGE_VISIBLE)
    ob = oraclebind.OracleOB()
    ob.set_size(w, h)
    ob.set_view(v)
    ob.set_max_triangles(1 << 30)
    ob.clear()
    ob.set_cull_mode(0)
    ob.draw_batch(model, vtx, 12, None, 0, 60)
    ob.build_hierarchy()
    assert np.array_equal(batch.depth(0), ob.depth())
    for a, b in zip(batch.mips(0), ob.mips()):
        assert np.array_equal(a, b)
    assert batch.tri_stats()[0, 1] == ob.num_triangles()
    # the drop-in object takes the same geometry through AddTriangles/DrawTriangles
    gob = capi.OcclusionBuffer(gpu_ctx)
    gob.SetSize(w, h)
    gob.SetView(v)
    gob.SetMaxTriangles(1 << 30)
    gob.Clear()
    gob.SetCullMode(0)
    gob.AddTriangles(model, vtx, 12, None, 0, 60)
    gob.DrawTriangles()
    assert np.array_equal(gob.GetBuffer(), ob.depth())
    for o in (gob, batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_config5_city_mesh_2048(gpu_ctx):
    """BASELINE config 5 shape: one 250k-triangle u32-indexed mesh (12-byte stride) into a 2048x2048 buffer (8 hierarchy levels: the last
    two come from the generic level kernel, 256 tiles, CTA-class triangles) and an occluded query over the config's 4M boxes, 1 view."""
    w = h = 2048
    vtx, idx = scenes.city_mesh(250_000, 500.0)
    sc = scenes.Scene(scenes.CONFIGS["C5"]["n"], 4, 500.0, seed=55)
    assert sc.n == 4_000_000
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, vtx, 12, idx)
    ident = np.array([[1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]], np.float32)
    occ = capi.Occluders(gpu_ctx, np.array([[-500, 0, -500, 500, 40, 500]], np.float32), ident, [mesh], cull_mode=capi.CULL_CCW)
    v = camera.make_view([20.0, 30.0, -480.0], 10.0, 12.0, 50.0, 1.0, 0.5, 1200.0)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, 1, sc.n)
    counts, offsets, ids = batch.cull_views(capi.pack_views([v]), capi.STAGE_VISIBLE)
    ob.set_view(v)
    ob.set_max_triangles(1 << 30)
    ob.clear()
    ob.set_cull_mode(capi.CULL_CCW)
    ob.draw_batch(ident[0], vtx, 12, idx, 0, idx.shape[0])
    ob.build_hierarchy()
    assert np.array_equal(batch.depth(0), ob.depth())
    gm, om = batch.mips(0), ob.mips()
    assert len(gm) == len(om) == 8
    for a, b in zip(gm, om):
        assert np.array_equal(a, b)
    assert batch.tri_stats()[0, 1] == ob.num_triangles()
    cand = otree.query(1, v.planes, ob, as_ids=False)
    vis = otree.order[otree.check_visibility(ob, cand)]
    assert counts[0, 0] == len(cand) and np.array_equal(ids[offsets[0]:offsets[1]], vis)
    assert 0 < len(vis) <= len(cand) < sc.n
    # the drop-in object draws the same mesh through AddTriangles / DrawTriangles (general path) and must agree as well
    gob = capi.OcclusionBuffer(gpu_ctx)
    gob.SetSize(w, h)
    gob.SetView(v)
    gob.SetMaxTriangles(1 << 30)
    gob.Clear()
    gob.SetCullMode(capi.CULL_CCW)
    gob.AddTriangles(ident[0], vtx, 12, idx, 0, idx.shape[0])
    gob.DrawTriangles()
    gob.BuildDepthHierarchy()
    assert np.array_equal(gob.GetBuffer(), ob.depth())
    assert gob.GetNumTriangles() == ob.num_triangles()
    for a, b in zip(gob.GetMips(), om):
        assert np.array_equal(a, b)
    for o in (gob, batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_config4_two_million_shadow_views(gpu_ctx):
    """BASELINE config 4: 4 ortho cascades + 64 point lights x 6 faces = 388 frustum-only ShadowCasterOctreeQuery views over 2M drawables.
    All 388 run on the GPU and every one of them is checked against the oracle's result sequence."""
    cfg = scenes.CONFIGS["C4"]
    sc = scenes.Scene(cfg["n"], 4, cfg["extent"], seed=4)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, _ = helpers.make_oracle(sc, flat, 8, 8)
    views = scenes.shadow_views(cfg["extent"], n_point_lights=64)
    assert len(views) == 388
    batch = capi.Batch(gpu_ctx, gs, None, 0, 0, len(views), sc.n)  # the 1000-unit cascade sees most of the scene
    batch.set_views(capi.pack_views(views, query_mode=capi.QUERY_SHADOWCASTER, use_occlusion=False))
    batch.run(capi.STAGE_VISIBLE)
    counts = batch.counts()
    offsets, ids = batch.packed(capacity=int(counts[:, 1].sum()))
    for i in range(388):
        want = otree.query(2, views[i].planes, None)
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], want), "view %d" % i
    assert counts[:, 1].sum() == offsets[-1] and counts[:4, 1].min() > 0
    for o in (batch, gs, tree):
        o.close()


@pytest.mark.parametrize("ortho,occl", [(False, True), (True, True), (False, False)])
def test_in_view_stage(gpu_ctx, ortho, occl):
    """U3D_STAGE_INVIEW: occlusion predicate + draw-distance test + per-view Z range (View.cpp:177-211, 926-951), batched."""
    w, h = 256, 160
    sc = helpers.small_scene(n=9000, k=128, extent=90.0, seed=23)
    rs = np.random.RandomState(6)
    draw_dist = np.where(rs.uniform(size=sc.n) < 0.5, rs.uniform(5.0, 150.0, size=sc.n), 0.0).astype(np.float32)
    sc.flags[::9] = 0x2
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    tree.set_draw_distances(draw_dist)
    flat = tree.flatten()
    gs = capi.GpuScene(gpu_ctx, octree=tree)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = []
    for i in range(10):
        pos = [rs.uniform(-60, 60), rs.uniform(5, 30) if ortho else 2.0, rs.uniform(-60, 60)]
        views.append(camera.make_view(pos, rs.uniform(0, 360), rs.uniform(20, 70) if ortho else rs.uniform(-10, 10), 45.0, w / h, 0.1, 300.0,
                                      ortho=ortho, ortho_size=rs.uniform(20, 100)))
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    mode = capi.QUERY_OCCLUDED if occl else capi.QUERY_FRUSTUM
    counts, offsets, ids = batch.cull_views(capi.pack_views(views, query_mode=mode, use_occlusion=occl), capi.STAGE_INVIEW)
    zr = batch.z_ranges()
    dropped = 0
    for i, v in enumerate(views):
        ob.draw_scene_view(sc, v)
        cand = otree.query(1 if occl else 0, v.planes, ob if occl else None, as_ids=False)
        pos_in, mm = otree.check_in_view(ob if occl else None, cand, v.view, v.position, ortho, draw_dist)
        assert counts[i, 0] == len(cand)
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], otree.order[pos_in]), "view %d" % i
        assert np.array_equal(zr[i], mm), (i, zr[i], mm)
        dropped += len(otree.check_visibility(ob if occl else None, cand)) - len(pos_in)
    assert dropped > 0  # the draw-distance test really removed something
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_incremental_box_updates(gpu_ctx):
    """Dirty-range maintenance: drawables that move inside their octant are refreshed in place (u3d_scene_update_boxes) and queries
    then equal those on a scene re-created from the same host octree, and the oracle on the moved boxes."""
    sc = helpers.small_scene(n=6000, k=4, extent=100.0, seed=41)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    rs = np.random.RandomState(2)
    boxes = sc.aabb.copy()
    moved_ids, stayed = [], 0
    for i in rs.choice(sc.n, size=1500, replace=False):
        boxes[i] += np.tile(rs.uniform(-0.3, 0.3, 3), 2).astype(np.float32)
        if tree.update(int(i), boxes[i]):
            stayed += 1
            moved_ids.append(int(i))
        else:
            boxes[i] = sc.aabb[i]
            tree.update(int(i), boxes[i])  # put it back; the layout check below decides whether the in-place path still applies
    flat2 = tree.flatten()
    assert stayed > 1000
    if not all(np.array_equal(flat[k], flat2[k]) for k in ("parent", "first", "count", "order")):
        pytest.skip("layout changed while restoring a moved drawable")
    gs.update_boxes(moved_ids, boxes[moved_ids])
    fresh = capi.GpuScene(gpu_ctx, octree=tree)
    otree, _ = helpers.make_oracle(scenes_with_boxes(sc, boxes), flat2, 8, 8)
    for v in scenes.agent_cameras(6, 100.0, 256, 160, seed=3):
        a, _ = gs.query(v.planes, None, mode=capi.QUERY_FRUSTUM)
        b, _ = fresh.query(v.planes, None, mode=capi.QUERY_FRUSTUM)
        assert np.array_equal(a, b)
        assert np.array_equal(a, otree.query(0, v.planes, None))
    for o in (fresh, gs, tree):
        o.close()


def test_update_boxes_with_duplicate_ids_keeps_the_last_box(gpu_ctx):
    """A drawable that moved twice before the upload appears twice in u3d_scene_update_boxes: the last box must win (as in the host
    tree), whatever order the device threads run in."""
    sc = helpers.small_scene(n=4000, k=4, extent=100.0, seed=43)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    rs = np.random.RandomState(4)
    boxes = sc.aabb.copy()
    ids, staged = [], []
    for i in rs.choice(sc.n, size=600, replace=False):
        first = boxes[i] + np.tile(rs.uniform(-0.2, 0.2, 3), 2).astype(np.float32)
        second = boxes[i] + np.tile(rs.uniform(-0.2, 0.2, 3), 2).astype(np.float32)
        if tree.update(int(i), first) and tree.update(int(i), second):
            ids += [int(i), int(i)]
            staged += [first, second]
            boxes[i] = second
        else:
            tree.update(int(i), boxes[i])
    flat2 = tree.flatten()
    if not all(np.array_equal(flat[k], flat2[k]) for k in ("parent", "first", "count", "order")):
        pytest.skip("layout changed while restoring a moved drawable")
    assert len(ids) > 600
    # interleave so that the two entries of a drawable are far apart in the launch
    perm = np.concatenate([np.arange(0, len(ids), 2), np.arange(1, len(ids), 2)])
    gs.update_boxes(np.array(ids)[perm], np.array(staged, np.float32)[perm])
    otree, _ = helpers.make_oracle(scenes_with_boxes(sc, boxes), flat2, 8, 8)
    for v in scenes.agent_cameras(6, 100.0, 256, 160, seed=5):
        a, _ = gs.query(v.planes, None, mode=capi.QUERY_FRUSTUM)
        assert np.array_equal(a, otree.query(0, v.planes, None))
    for o in (gs, tree):
        o.close()


def test_batch_follows_the_scene_after_structural_changes(gpu_ctx):
    """u3d_batch_set_scene: after inserts / removals / moves across octants the flattened scene is a new object (more octants, other
    order); a batch created before must run against it, not against the destroyed one."""
    sc = helpers.small_scene(n=3000, k=4, extent=60.0, seed=45)
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    half = sc.n // 2
    tree.add(sc.aabb[:half], sc.flags[:half], sc.view_mask[:half], sc.occludee[:half], sc.cast_shadows[:half])
    gs = capi.GpuScene(gpu_ctx, octree=tree)
    views = scenes.agent_cameras(20, 60.0, 256, 160, seed=6)
    packed = capi.pack_views(views, query_mode=capi.QUERY_FRUSTUM)
    batch = capi.Batch(gpu_ctx, gs, None, 2, 2, len(views), sc.n)
    c0, o0, i0 = batch.cull_views(packed, capi.STAGE_QUERY)
    # structural change: the other half arrives, some drawables leave, some jump across the world
    tree.add(sc.aabb[half:], sc.flags[half:], sc.view_mask[half:], sc.occludee[half:], sc.cast_shadows[half:])
    boxes = sc.aabb.copy()
    for i in range(0, half, 7):
        tree.remove(i)
    for i in range(3, half, 11):
        if i % 7:
            boxes[i] = boxes[i] + np.tile(np.float32([35.0, 0.0, -28.0]), 2)
            tree.update(i, boxes[i])
    flat = tree.flatten()
    gs2 = capi.GpuScene(gpu_ctx, octree=tree)
    gs.close()  # the old scene is gone, as after GpuOctree::Sync
    batch.set_scene(gs2)
    c1, o1, i1 = batch.cull_views(packed, capi.STAGE_QUERY)
    otree, _ = helpers.make_oracle(scenes_with_boxes(sc, boxes), flat, 8, 8)
    for k, v in enumerate(views):
        want = otree.query(0, v.planes, None)
        assert np.array_equal(i1[o1[k]:o1[k + 1]], want), "view %d" % k
    assert not np.array_equal(c0, c1)
    for o in (batch, gs2, tree):
        o.close()


def scenes_with_boxes(sc, boxes):
    import copy
    c = copy.copy(sc)
    c.aabb = boxes
    return c


def test_sphere_box_point_queries(gpu_ctx):
    """U3D_QUERY_SPHERE / BOX / POINT through u3d_octree_query and through the batched entry point."""
    sc = helpers.small_scene(n=20000, k=4, extent=150.0, seed=29)
    sc.flags[::7] = 0x2
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, _ = helpers.make_oracle(sc, flat, 8, 8)
    rs = np.random.RandomState(3)
    from test_oracle_vs_reference import shape_cases
    cases = shape_cases(rs, 150.0, n=10)
    total = 0
    for shape, params in cases:
        want = otree.query(shape, params, None)
        got, qc = gs.query(params, None, mode=shape)
        assert qc == len(want) and np.array_equal(got, want), (shape, params)
        total += len(want)
    assert total > 0
    # batched: one "view" per query
    packed = np.zeros(len(cases), capi.VIEW_DTYPE)
    for i, (shape, params) in enumerate(cases):
        packed[i]["planes"][:len(params)] = params
        packed[i]["query_mode"] = shape
    packed["drawable_flags"] = 0xFF
    packed["view_mask"] = 0xFFFFFFFF
    batch = capi.Batch(gpu_ctx, gs, None, 0, 0, len(cases), sc.n)
    counts, offsets, ids = batch.cull_views(packed, capi.STAGE_QUERY)
    for i, (shape, params) in enumerate(cases):
        assert np.array_equal(ids[offsets[i]:offsets[i + 1]], otree.query(shape, params, None))
    for o in (batch, gs, tree):
        o.close()


def test_pipelined_batches_from_two_threads(gpu_ctx):
    """Two batches over one scene, each driven by its own host thread on its own stream with the pipelined hint set: every call returns the
    same visible sets as the oracle, whatever the overlap on the device."""
    import threading

    import torch

    w = h = 128
    nviews = 48
    sc = helpers.small_scene(n=12000, k=120, extent=100.0, seed=91)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    sets = [scenes.agent_cameras(nviews, 100.0, w, h, seed=5), helpers.stress_views(100.0, w, h, nviews, seed=6)]
    want = [[oracle_view(sc, otree, ob, v)[2] for v in vs] for vs in sets]
    batches = [capi.Batch(gpu_ctx, gs, occ, w, h, nviews, sc.n) for _ in sets]
    streams = [torch.cuda.Stream() for _ in sets]
    errs = []

    def worker(t):
        try:
            batches[t].set_pipelined(True)
            packed = capi.pack_views(sets[t])
            for _ in range(6):
                counts, offsets, ids = batches[t].cull_views(packed, capi.STAGE_VISIBLE, streams[t].cuda_stream)
                for i in range(nviews):
                    assert counts[i, 1] == len(want[t][i])
                    assert np.array_equal(ids[offsets[i]:offsets[i + 1]], want[t][i]), "thread %d view %d" % (t, i)
        except BaseException as e:  # noqa: BLE001 - re-raised below
            errs.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(len(sets))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errs:
        raise errs[0]
    for o in (*batches, occ, mesh, gs, tree):
        o.close()
    ob.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_memory_result_exchange(gpu_ctx, world):
    """The multi-GPU exchange (u3d_gather_*: fused pack + stores into every rank's receive block + arrival flags) with `world` logical ranks
    living in this process on one device: after push / wait every rank's block holds every rank's counts, packed offsets and visible ids,
    equal to the oracle's sets in global view order; repeated epochs (with release) and a too-small ids region behave as documented."""
    import torch

    w = h = 128
    nviews = 31  # not a multiple of world: the last rank slots stay short
    sc = helpers.small_scene(n=9000, k=100, extent=100.0, seed=17)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    views = scenes.agent_cameras(nviews, 100.0, w, h, seed=8)
    want = [oracle_view(sc, otree, ob, v)[2] for v in views]
    per = (nviews + world - 1) // world
    from urho3d_b200 import sharding
    shards = [sharding.shard_views(nviews, r, world) for r in range(world)]
    batches, gathers, streams = [], [], []
    cap = sum(len(x) for x in want) + 16
    for r in range(world):
        b = capi.Batch(gpu_ctx, gs, occ, w, h, per, sc.n)
        b.set_views(capi.pack_views([views[i] for i in shards[r]]))
        batches.append(b)
        gathers.append(capi.Gather(gpu_ctx, world, r, per, cap))
        streams.append(torch.cuda.Stream())
    for g in gathers:
        g.connect_local(gathers)
    for epoch in range(3):
        for r in range(world):
            batches[r].run(capi.STAGE_VISIBLE, streams[r].cuda_stream)
            gathers[r].push(batches[r], streams[r].cuda_stream)
        for r in range(world):
            gathers[r].wait(streams[r].cuda_stream)
        for r in range(world):
            counts, offsets, ids = gathers[r].read(streams[r].cuda_stream)
            for v in range(nviews):
                src, slot = v % world, v // world
                assert counts[src, slot, 1] == len(want[v])
                o = offsets[src, slot]
                assert np.array_equal(ids[src, o:o + len(want[v])], want[v]), "epoch %d rank %d view %d" % (epoch, r, v)
            assert offsets[0, len(shards[0])] == sum(len(want[v]) for v in shards[0])
        for r in range(world):
            gathers[r].release(streams[r].cuda_stream)
    # a rank that pushes fewer views than before (uneven shards, a last partial batch): its unused slots read as empty everywhere,
    # not as the previous epoch's counts; then full again
    short = min(2, len(shards[0]))
    for nfirst in (short, len(shards[0])):
        batches[0].set_views(capi.pack_views([views[i] for i in shards[0][:nfirst]]))
        for r in range(world):
            batches[r].run(capi.STAGE_VISIBLE, streams[r].cuda_stream)
            gathers[r].push(batches[r], streams[r].cuda_stream)
        for r in range(world):
            gathers[r].wait(streams[r].cuda_stream)
        for r in range(world):
            counts, offsets, ids = gathers[r].read(streams[r].cuda_stream)
            tot0 = sum(len(want[v]) for v in shards[0][:nfirst])
            assert np.all(counts[0, nfirst:] == 0), "stale counts in unused slots"
            assert np.all(offsets[0, nfirst:] == tot0)
            for slot, v in enumerate(shards[0][:nfirst]):
                o = offsets[0, slot]
                assert counts[0, slot, 1] == len(want[v]) and np.array_equal(ids[0, o:o + len(want[v])], want[v])
            for src in range(1, world):
                for slot, v in enumerate(shards[src]):
                    o = offsets[src, slot]
                    assert np.array_equal(ids[src, o:o + len(want[v])], want[v])
        for r in range(world):
            gathers[r].release(streams[r].cuda_stream)
    # capacity: a region that cannot hold a rank's ids is reported, not silently truncated
    small = [capi.Gather(gpu_ctx, world, r, per, 8) for r in range(world)]
    for g in small:
        g.connect_local(small)
    for r in range(world):
        small[r].push(batches[r], streams[r].cuda_stream)
    for r in range(world):
        small[r].wait(streams[r].cuda_stream)
    with pytest.raises(capi.U3DError) as e:
        small[0].status(streams[0].cuda_stream)
    assert e.value.code == 4
    small[0].status(streams[0].cuda_stream)  # reported once, then cleared: a later status of the same object is clean
    for o in (*small, *gathers, *batches, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_raycasts_match_oracle(gpu_ctx):
    """Octree::Raycast / RaycastSingle (RAY_AABB level) for a batch of rays: ids, bit-equal distances and positions in the reference's result
    order (ties included), with filters, finite maxDistance, rays from inside boxes, a ray that misses the octree, and the capacity error."""
    sc = helpers.small_scene(n=20000, k=8, extent=120.0, seed=31)
    sc.aabb[300:400] = sc.aabb[500:600]  # stacked duplicates: equal distances
    rs = np.random.RandomState(6)
    sc.flags[rs.uniform(size=sc.n) < 0.2] = 2
    sc.view_mask[rs.uniform(size=sc.n) < 0.2] = 0x2
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, 8, 8)
    rays = helpers.random_rays(120.0, 200, seed=12, stacked_boxes=sc.aabb[300:400])
    rays.append((np.array([0.0, 5000.0, 0.0], np.float32), np.array([0.0, 1.0, 0.0], np.float32), np.float32(np.inf)))  # starts above, points away
    rays7 = np.array([list(o) + list(d) + [m] for o, d, m in rays], np.float32)
    total = 0
    for flags, mask in ((0xFF, 0xFFFFFFFF), (0x1, 0x1)):
        for single in (False, True):
            got = gs.raycast(rays7, flags, mask, single, capacity=8192)
            for i, (o, d, m) in enumerate(rays):
                wid, wdist = otree.raycast(o, d, m, flags, mask, single)
                h = got[i]
                assert np.array_equal(h["drawable"], wid.astype(np.uint32)), "ray %d single %d" % (i, single)
                assert np.array_equal(h["distance"].view(np.uint32), wdist.view(np.uint32))
                pos = (o[None, :] + wdist[:, None] * d[None, :]).astype(np.float32)
                assert np.array_equal(h["position"], pos) and np.array_equal(h["normal"], np.broadcast_to(-d, pos.shape))
                total += len(wid)
    assert total > 1000 and len(got[-1]) == 0
    with pytest.raises(capi.U3DError) as e:
        gs.raycast(rays7[:8], capacity=1)
    assert e.value.code == 4
    for o in (gs, tree):
        o.close()
    ob.close()


@pytest.mark.parametrize("ortho", [False, True])
def test_occluder_heuristics_and_budget(gpu_ctx, ortho):
    """SURVEY 8f row 3 on the device: View::UpdateOccluders' selection + DrawOccluders' triangle budget per view.  The set of submitted
    occluders, the submitted-triangle count, the depth plane and the visible set equal the oracle's, including budgets that cut between
    occluders with EQUAL sort keys (there the reference's unstable Sort decides, and the device replays it)."""
    w = h = 128
    sc = helpers.small_scene(n=6000, k=300, extent=90.0, seed=41)
    sc.occ_aabb[50:90] = sc.occ_aabb[100:140]
    sc.occ_model[50:90] = sc.occ_model[100:140]
    rs = np.random.RandomState(3)
    draw_dist = np.where(rs.uniform(size=sc.k) < 0.3, rs.uniform(20.0, 120.0, size=sc.k), 0.0).astype(np.float32)
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    occ.set_draw_distances(draw_dist)
    tri = sc.mesh_idx.shape[0] // 3
    views = []
    for i in range(20):
        pos = [rs.uniform(-90, 90), rs.uniform(1.0, 30.0), rs.uniform(-90, 90)]
        if i % 5 == 0:
            b = sc.occ_aabb[rs.randint(sc.k)]
            pos = list((b[:3] + b[3:]) * 0.5)
        views.append(camera.make_view(pos, rs.uniform(0, 360), rs.uniform(-30, 10), 45.0, 1.0, 0.1, rs.choice([150.0, 400.0]), ortho=ortho,
                                      ortho_size=rs.choice([20.0, 60.0])))
    cam3 = np.array([[v.half_view_size, v.ortho_size, v.far] for v in views], np.float32)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    batch.set_views(capi.pack_views(views))
    tie_cuts = 0
    # a budget per round; the last rounds put the cut of view `target` exactly between two equal keys
    budgets = [(0.025, 5000), (0.1, 30 * tri), (0.0, 7 * tri + 5), (0.025, 1 << 30)]
    for target in range(len(views)):
        oidx, okeys, _ = oraclebind.select_occluders(sc, views[target], 0.0, 1 << 30, draw_dist)
        ties = np.nonzero(okeys[1:] == okeys[:-1])[0]
        if len(ties):
            budgets.append((0.0, int(ties[0]) * tri + tri - 1))
    assert len(budgets) > 6
    for thr, budget in budgets:
        batch.set_occluder_policy(thr, budget, cam3)
        batch.run(capi.STAGE_VISIBLE)
        counts = batch.counts()
        stats = batch.tri_stats()
        offsets, ids = batch.packed(capacity=int(counts[:, 1].sum()))
        for i, v in enumerate(views):
            oidx, okeys, nsub = oraclebind.select_occluders(sc, v, thr, budget, draw_dist)
            got = batch.occluder_selection(i)
            assert sorted(got.tolist()) == sorted(oidx[:nsub].tolist()), "view %d thr %g budget %d" % (i, thr, budget)
            assert stats[i, 0] == nsub * tri
            if nsub < len(oidx) and okeys[nsub - 1] == okeys[nsub]:
                tie_cuts += 1
            ob.draw_scene_view_selected(sc, v, oidx[:nsub], budget)
            assert np.array_equal(batch.depth(i), ob.depth())
            cand = otree.query(1, v.planes, ob, as_ids=False)
            assert np.array_equal(ids[offsets[i]:offsets[i + 1]], otree.order[otree.check_visibility(ob, cand)])
    assert tie_cuts > 0
    # back to the default rule
    batch.set_occluder_policy(0, 0, None, enable=False)
    batch.run(capi.STAGE_VISIBLE)
    ob.draw_scene_view(sc, views[0])
    assert np.array_equal(batch.depth(0), ob.depth())
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_full_size_properties(gpu_ctx):
    """BASELINE config 3 at full scene size (1M drawables, 4000 occluders, 256x256) over 320 views, checked through properties that need no
    oracle pass: (a) every launch shape / queue size / walk variant of the cull kernel and a shuffled view order give identical sequences;
    (b) the visible sequence is an order-preserving subsequence of the occluded query result, which is one of the frustum query result;
    (c) without occluders the occluded query and the predicate pass everything the frustum query finds; (d) the heuristic occluder policy with
    an unlimited budget and zero threshold draws the same buffers as the default rule; (e) a second run over the same slot changes nothing."""
    cfg = scenes.CONFIGS["C3"]
    sc = scenes.Scene(cfg["n"], cfg["k"], cfg["extent"])
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    none = capi.Occluders(gpu_ctx, sc.occ_aabb[:1] + 1.0e6, sc.occ_model[:1], [mesh], cull_mode=sc.cull_mode)  # one occluder far outside
    V = 320
    w = h = 256
    views = scenes.agent_cameras(V, cfg["extent"], w, h, seed=77)
    packed = capi.pack_views(views)

    def run(batch, pk, stage=capi.STAGE_VISIBLE):
        batch.set_views(pk)
        batch.run(stage)
        c = batch.counts()
        off, ids = batch.packed(capacity=int(c[:, 1].sum()))
        return c, off, ids

    cap = 450000  # a 45-degree frustum reaching across the 1000 x 1000 region holds up to ~0.4M of the 1M drawables
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, V, cap)
    c0, off0, ids0 = run(batch, packed)
    depth0 = batch.depth(5).copy()
    assert c0[:, 1].min() > 0 and (c0[:, 1] <= c0[:, 0]).all() and (c0[:, 1] < c0[:, 0]).any()
    # (e) idempotence
    c1, off1, ids1 = run(batch, packed)
    assert np.array_equal(c0, c1) and np.array_equal(ids0, ids1)
    # (a) launch shapes and tuning variants
    for opt, val, restore in (("cull_ring", 2048, 4096), ("cull_ring", 8192, 4096), ("walk_pool", 1, 0), ("heavy_extent", 24, 96)):
        capi.debug_set_option(opt, val)
        try:
            batch.set_pipelined(True)
            c2, off2, ids2 = run(batch, packed)
        finally:
            capi.debug_set_option(opt, restore)
            batch.set_pipelined(False)
        assert np.array_equal(c0, c2) and np.array_equal(ids0, ids2), opt
    batch.set_pipelined(True)   # 320 views: three-CTA shape instead of the one-CTA-per-SM shape
    c2, off2, ids2 = run(batch, packed)
    batch.set_pipelined(False)
    assert np.array_equal(c0, c2) and np.array_equal(ids0, ids2)
    perm = np.random.RandomState(1).permutation(V)
    c3, off3, ids3 = run(batch, packed[perm])
    assert np.array_equal(c3, c0[perm])
    for j in (0, 17, V - 1):
        assert np.array_equal(ids3[off3[j]:off3[j + 1]], ids0[off0[perm[j]]:off0[perm[j] + 1]])
    # (b) subsequence chain: visible <= occluded query <= frustum query
    cq, offq, idsq = run(batch, packed, capi.STAGE_QUERY)
    cf, offf, idsf = run(batch, capi.pack_views(views, query_mode=capi.QUERY_FRUSTUM, use_occlusion=False), capi.STAGE_QUERY)

    def is_subsequence(a, b):
        index_of = {int(x): i for i, x in enumerate(b)}
        last = -1
        for x in a:
            i = index_of.get(int(x), -1)
            if i <= last:
                return False
            last = i
        return True

    assert np.array_equal(cq[:, 0], c0[:, 0]) and (cq[:, 1] == cq[:, 0]).all() and (cf[:, 0] >= cq[:, 0]).all()
    for j in range(0, V, 16):
        vis = ids0[off0[j]:off0[j + 1]]
        qry = idsq[offq[j]:offq[j + 1]]
        fru = idsf[offf[j]:offf[j + 1]]
        assert is_subsequence(vis, qry) and is_subsequence(qry, fru), j
    # (c) no occluder in sight: nothing is culled by occlusion
    b2 = capi.Batch(gpu_ctx, gs, none, w, h, V, cap)
    cn, offn, idsn = run(b2, packed)
    assert np.array_equal(cn[:, 0], cf[:, 0]) and np.array_equal(cn[:, 1], cf[:, 1]) and np.array_equal(idsn, idsf)
    assert (b2.depth(3) == 16777216).all()
    b2.close()
    # (d) heuristic selection without limits == default rule
    batch.set_views(packed)
    cam3 = np.array([[v.half_view_size, v.ortho_size, v.far] for v in views], np.float32)
    batch.set_occluder_policy(0.0, 1 << 30, cam3)
    batch.run(capi.STAGE_VISIBLE)
    c4 = batch.counts()
    off4, ids4 = batch.packed(capacity=int(c4[:, 1].sum()))
    assert np.array_equal(c4, c0) and np.array_equal(ids4, ids0) and np.array_equal(batch.depth(5), depth0)
    for o in (batch, none, occ, mesh, gs, tree):
        o.close()


def test_zone_occluder_query(gpu_ctx):
    """U3D_QUERY_ZONE_OCCLUDER (ZoneOccluderOctreeQuery, View.cpp:87-113) through u3d_octree_query and the batched entry point, occluder flags set
    on the host mirror (u3d_octree_set_occluder_flags) and in place on the device (u3d_scene_set_occluder_flags)."""
    sc = helpers.small_scene(n=15000, k=8, extent=100.0, seed=61)
    rs = np.random.RandomState(8)
    sc.flags[:] = rs.choice([1, 2, 4, 5, 8], size=sc.n, p=[0.6, 0.1, 0.1, 0.1, 0.1]).astype(np.uint8)
    sc.view_mask[rs.uniform(size=sc.n) < 0.2] = 0x2
    occluder = (rs.uniform(size=sc.n) < 0.3).astype(np.uint8)
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    tree.set_occluder_flags(occluder)
    flat = tree.flatten()
    gs = capi.GpuScene(gpu_ctx, octree=tree)
    otree, ob = helpers.make_oracle(sc, flat, 8, 8)
    views = scenes.agent_cameras(10, 100.0, 64, 64, seed=3) + helpers.stress_views(100.0, 64, 64, 6, seed=4)
    total = 0
    for mask in (0xFFFFFFFF, 0x1):
        batch = capi.Batch(gpu_ctx, gs, None, 0, 0, len(views), sc.n)
        batch.set_views(capi.pack_views(views, view_mask=mask, query_mode=capi.QUERY_ZONE_OCCLUDER, use_occlusion=False))
        batch.run(capi.STAGE_QUERY)
        counts = batch.counts()
        offsets, ids = batch.packed(capacity=int(counts[:, 1].sum()))
        for i, v in enumerate(views):
            want = otree.query_zone_occluder(v.planes, occluder, mask)
            assert np.array_equal(ids[offsets[i]:offsets[i + 1]], want), "view %d" % i
            single, _ = gs.query(v.planes, None, 0xFF, mask, capi.QUERY_ZONE_OCCLUDER, capi.STAGE_QUERY)
            assert np.array_equal(single, want)
            total += len(want)
        batch.close()
    assert total > 1000
    # flip the flags in place on the device
    occluder2 = (1 - occluder).astype(np.uint8)
    gs.set_occluder_flags(occluder2)
    want = otree.query_zone_occluder(views[0].planes, occluder2, 0xFFFFFFFF)
    got, _ = gs.query(views[0].planes, None, 0xFF, 0xFFFFFFFF, capi.QUERY_ZONE_OCCLUDER, capi.STAGE_QUERY)
    assert np.array_equal(got, want) and len(want) > 0
    # the other queries do not see the occluder bit
    f1, _ = gs.query(views[0].planes, None, 0xFF, 0xFFFFFFFF, capi.QUERY_FRUSTUM, capi.STAGE_QUERY)
    assert np.array_equal(f1, otree.query(0, views[0].planes, None))
    for o in (gs, tree):
        o.close()
    ob.close()


def test_process_shadow_casters(gpu_ctx):
    """View::ProcessShadowCasters + IsShadowCasterVisible on the device, one split per batch view (directional cascades, spot
    lights, point-light cube faces), over the lists a frustum query emitted; the filtered sequences equal the oracle's (which is pinned to the
    reference's math in tests/test_oracle_vs_reference.py).  The split inputs come from the reference's own Camera / Frustum code."""
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref/libu3dref.so not built")
    sc = helpers.small_scene(n=20000, k=8, extent=100.0, seed=71)
    rs = np.random.RandomState(9)
    shadow_mask = np.where(rs.uniform(size=sc.n) < 0.15, 0x2, 0xFFFFFFFF).astype(np.uint32)
    shadow_dist = np.where(rs.uniform(size=sc.n) < 0.3, rs.uniform(10.0, 80.0, size=sc.n), 0.0).astype(np.float32)
    draw_dist = np.where(rs.uniform(size=sc.n) < 0.3, rs.uniform(10.0, 80.0, size=sc.n), 0.0).astype(np.float32)
    in_view = (rs.uniform(size=sc.n) < 0.3).astype(np.uint8)
    ref = refbind.Ref(0)
    scen = helpers.shadow_scenarios(ref, 100.0, 36, seed=6, shared_main=True)
    ref.close()
    assert len(scen) >= 24 and {s["light_type"] for s in scen} == {0, 1, 2}
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    tree.set_draw_distances(draw_dist)
    flat = tree.flatten()
    gs = capi.GpuScene(gpu_ctx, octree=tree)
    gs.set_shadow_params(shadow_mask, shadow_dist)
    otree, ob = helpers.make_oracle(sc, flat, 8, 8)
    views = []
    for s in scen:
        planes = s["split"]["shadow_planes"].reshape(6, 4)
        views.append(camera.CameraView(s["split"]["light_view"].reshape(3, 4), np.eye(4, dtype=np.float32), planes, 0.1, s["split"]["shadow_far"]))
    batch = capi.Batch(gpu_ctx, gs, None, 0, 0, len(views), sc.n)
    kept = dropped = 0
    for light_mask in (0xFFFFFFFF, 0x1):
        splits = np.zeros(len(scen), capi.SHADOW_SPLIT_DTYPE)
        for i, s in enumerate(scen):
            sp = s["split"]
            splits[i]["light_view"] = sp["light_view"]
            splits[i]["shadow_camera_frustum"] = sp["shadow_planes"]
            splits[i]["light_view_frustum"] = sp["lvf_planes"]
            splits[i]["light_view_frustum_max_z"] = sp["lvf_max_z"]
            splits[i]["shadow_far_clip"] = sp["shadow_far"]
            splits[i]["shadow_orthographic"] = int(sp["shadow_ortho"])
            splits[i]["light_type"] = s["light_type"]
            splits[i]["light_mask"] = light_mask
        splits[3]["skip"] = 1
        batch.set_views(capi.pack_views(views, query_mode=capi.QUERY_FRUSTUM, use_occlusion=False))
        batch.run(capi.STAGE_QUERY)
        before = batch.counts()
        batch.process_shadow_casters(splits, scen[0]["main_view"], scen[0]["main_pos"], False, in_view)
        counts = batch.counts()
        assert np.array_equal(counts[:, 0], before[:, 0])  # the query result size stays, the emitted list shrinks
        offsets, ids = batch.packed(capacity=int(counts[:, 1].sum()))
        for i, s in enumerate(scen):
            cand = otree.query(0, s["split"]["shadow_planes"], None)
            assert before[i, 1] == len(cand)
            want = oraclebind.process_shadow_casters(cand, sc.aabb, sc.cast_shadows, s["split"], s["light_type"], light_mask, shadow_mask,
                                                     shadow_dist, draw_dist, in_view, s["main_view"], s["main_pos"])
            if i == 3:
                want = want[:0]
            assert np.array_equal(ids[offsets[i]:offsets[i + 1]], want), "split %d type %d" % (i, s["light_type"])
            kept += len(want)
            dropped += len(cand) - len(want)
    assert kept > 2000 and dropped > 2000
    for o in (batch, gs, tree):
        o.close()
    ob.close()


@pytest.mark.parametrize("w,h", [(128, 128), (256, 160)])
def test_serial_occluder_branch(gpu_ctx, w, h):
    """Occluder policy 2 = View::DrawOccluders' NON-threaded branch (the engine default) per view on the device: activeOccluders_, numTriangles_,
    the depth plane, every mip and the visible sequence equal the oracle's serial loop (pinned to the reference's own OcclusionBuffer), with
    budgets that stop the loop, duplicated occluders (equal sort keys: the list ORDER matters here) and stress views that clip."""
    sc = helpers.small_scene(n=5000, k=300, extent=90.0, seed=43)
    sc.occ_aabb[50:90] = sc.occ_aabb[100:140]
    sc.occ_model[50:90] = sc.occ_model[100:140]
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob = helpers.make_oracle(sc, flat, w, h)
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    rs = np.random.RandomState(5)
    tri = sc.mesh_idx.shape[0] // 3
    views = []
    for i in range(14):
        pos = [rs.uniform(-90, 90), rs.uniform(1.0, 20.0), rs.uniform(-90, 90)]
        views.append(camera.make_view(pos, rs.uniform(0, 360), rs.uniform(-25, 10), 45.0, float(w) / h, 0.1, rs.choice([150.0, 400.0])))
    for v in helpers.stress_views(90.0, w, h, 6, seed=3):
        v.half_view_size = float(np.float32(np.tan(np.float32(0.4))))
        views.append(v)
    cam3 = np.array([[v.half_view_size, v.ortho_size, v.far] for v in views], np.float32)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    batch.set_views(capi.pack_views(views))
    skipped = cut = 0
    for thr, budget in ((0.025, 5000), (0.0, 40 * tri), (0.025, 600), (0.025, 1 << 30)):
        batch.set_occluder_policy(thr, budget, cam3, enable=2)
        batch.run(capi.STAGE_VISIBLE)
        counts = batch.counts()
        stats = batch.tri_stats()
        active = batch.active_occluders()
        offsets, ids = batch.packed(capacity=int(counts[:, 1].sum()))
        for i, v in enumerate(views):
            oidx, _, _ = oraclebind.select_occluders(sc, v, thr, budget)
            want_active, want_tris = ob.draw_scene_view_serial(sc, v, oidx, budget)
            assert active[i] == want_active and stats[i, 1] == want_tris, "view %d thr %g budget %d" % (i, thr, budget)
            assert np.array_equal(batch.depth(i), ob.depth())
            for a, b in zip(batch.mips(i), ob.mips()):
                assert np.array_equal(a, b)
            cand = otree.query(1, v.planes, ob, as_ids=False)
            assert counts[i, 0] == len(cand)
            assert np.array_equal(ids[offsets[i]:offsets[i + 1]], otree.order[otree.check_visibility(ob, cand)])
            skipped += want_active < len(oidx)
            cut += budget < (1 << 30) and want_tris > budget
    assert skipped > 10 and cut > 3
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()


def test_empty_and_ragged_inputs(gpu_ctx):
    """Edge cases: an octree without drawables, a view that sees nothing, a batch called with zero views, an occlusion buffer drawn with no
    batches / with a triangle count that is not a multiple of three, zero rays, and a batch whose views have wildly different result sizes
    (ragged packed output).  Everything is compared with the oracle where the reference defines a result."""
    w, h = 64, 30  # odd mip heights on the way down
    # --- scene without drawables: only the root octant
    empty = capi.Octree()
    es = capi.GpuScene(gpu_ctx, octree=empty)
    assert es.num_drawables == 0 and es.num_octants == 1
    v0 = camera.make_view([0.0, 2.0, 0.0], 30.0, -5.0, 45.0, float(w) / h, 0.1, 300.0)
    ids, qn = es.query(v0.planes)
    assert len(ids) == 0 and qn == 0
    assert [len(x) for x in es.raycast(np.array([[0, 5, 0, 0, -1, 0, np.inf]], np.float32))] == [0]
    assert es.raycast(np.zeros((0, 7), np.float32)) == []
    be = capi.Batch(gpu_ctx, es, None, 0, 0, 4, 16)
    be.set_views(capi.pack_views([v0, v0], query_mode=capi.QUERY_FRUSTUM, use_occlusion=False))
    be.run(capi.STAGE_VISIBLE)
    assert (be.counts() == 0).all()
    off, pk = be.packed()
    assert off.tolist() == [0, 0, 0] and len(pk) == 0
    be.set_views(capi.pack_views([v0], query_mode=capi.QUERY_FRUSTUM, use_occlusion=False)[:0])   # zero views: a no-op
    be.run(capi.STAGE_VISIBLE)
    assert be.counts().shape == (0, 2)
    be.close()
    es.close()
    empty.close()

    # --- occlusion buffer: nothing drawn, then a draw call whose index count is 7 (two triangles + one stray index)
    sc = helpers.small_scene(n=2000, k=24, extent=60.0, seed=13)
    ob = oraclebind.OracleOB()
    ob.set_size(w, h)
    gob = capi.OcclusionBuffer(gpu_ctx)
    assert gob.SetSize(w, h)
    gob.SetView(v0)
    ob.set_view(v0)
    gob.Clear()
    ob.clear()
    gob.DrawTriangles()
    gob.BuildDepthHierarchy()
    ob.build_hierarchy()
    assert np.array_equal(gob.GetBuffer(), ob.depth()) and (ob.depth() == 16777216).all()
    for a, b in zip(gob.GetMips(), ob.mips()):
        assert np.array_equal(a, b)
    assert gob.IsVisible(sc.aabb[0]) == bool(ob.is_visible(sc.aabb[0]))
    gob.SetCullMode(capi.CULL_NONE)
    ob.set_cull_mode(capi.CULL_NONE)
    model = np.array([6, 0, 0, 0, 0, 6, 0, 2, 0, 0, 1, 20], np.float32)
    # DrawBatch (OB.cpp:464-541): an indexed count of 7 draws three triangles (the loop runs while indices < indicesEnd, reading indices 7 and 8
    # of the buffer too), numTriangles_ counts 7 / 3 = 2 submitted
    r_g = gob.AddTriangles(model, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx, 0, 7)
    r_o = ob.draw_batch(model, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx, 0, 7)
    gob.DrawTriangles()
    assert bool(r_g) == bool(r_o) and np.array_equal(gob.GetBuffer(), ob.depth()) and gob.GetNumTriangles() == ob.num_triangles()
    # ... and a non-indexed count of 8 draws two triangles (index + 2 < drawCount)
    flat_vtx = np.ascontiguousarray(sc.mesh_vtx.view(np.uint8).reshape(-1, sc.mesh_stride)[sc.mesh_idx[:9]]).reshape(-1)
    r_g = gob.AddTriangles(model, flat_vtx, sc.mesh_stride, None, 0, 8)
    r_o = ob.draw_batch(model, flat_vtx, sc.mesh_stride, None, 0, 8)
    gob.DrawTriangles()
    assert bool(r_g) == bool(r_o) and np.array_equal(gob.GetBuffer(), ob.depth()) and gob.GetNumTriangles() == ob.num_triangles()
    gob.close()
    with pytest.raises(capi.U3DError) as e:  # the batched occluder table refuses what it would draw differently
        m0 = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
        capi.Occluders(gpu_ctx, sc.occ_aabb[:1], sc.occ_model[:1], [m0], start=np.zeros(1, np.uint32), count=np.full(1, 7, np.uint32))
    assert e.value.code == 6

    # --- ragged results: one camera sees most of the scene, others see nothing
    tree, flat, gs = gpu_scene(gpu_ctx, sc)
    otree, ob2 = helpers.make_oracle(sc, flat, w, h)
    views = [camera.make_view([0.0, 80.0, 0.0], 0.0, 89.0, 100.0, float(w) / h, 0.1, 500.0),      # looks straight down on everything
             camera.make_view([0.0, 2.0, 0.0], 0.0, -89.0, 20.0, float(w) / h, 0.1, 300.0),        # looks at the sky
             camera.make_view([5000.0, 2.0, 5000.0], 45.0, 0.0, 45.0, float(w) / h, 0.1, 100.0),   # far outside the octree
             camera.make_view([10.0, 2.0, -30.0], 200.0, -3.0, 45.0, float(w) / h, 0.1, 300.0)]
    mesh = capi.Mesh(gpu_ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(gpu_ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    batch = capi.Batch(gpu_ctx, gs, occ, w, h, len(views), sc.n)
    counts, offsets, ids = batch.cull_views(capi.pack_views(views), capi.STAGE_VISIBLE)
    sizes = []
    for i, v in enumerate(views):
        sub, cand, vis = oracle_view(sc, otree, ob2, v)
        assert np.array_equal(batch.depth(i), ob2.depth())
        assert counts[i, 0] == len(cand) and np.array_equal(ids[offsets[i]:offsets[i + 1]], vis), "view %d" % i
        sizes.append(len(vis))
    assert max(sizes) > 500 and min(sizes) == 0
    for o in (batch, occ, mesh, gs, tree):
        o.close()
    ob.close()
    ob2.close()
