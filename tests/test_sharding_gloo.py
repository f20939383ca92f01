"""world_size-2 gloo test of the multi-GPU host logic (view sharding + gather of counts and visible sets), CPU only.
Each rank computes its shard of views with the C oracle standing in for its GPU; the gathered result must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from urho3d_b200 import scenes, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _visible_sets(sc, flat, views, w, h):
    tree, ob = helpers.make_oracle(sc, flat, w, h)
    out = []
    for v in views:
        ob.draw_scene_view(sc, v)
        cand = tree.query(1, v.planes, ob, as_ids=False)
        vis = tree.order[tree.check_visibility(ob, cand)]
        out.append((len(cand), vis.astype(np.int32)))
    ob.close()
    return out


def _worker(rank, world, port, n_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 128, 80
    sc = helpers.small_scene(n=3000, k=96, extent=80.0, seed=7)
    _, flat = helpers.host_flatten(sc)
    views = scenes.agent_cameras(n_views, 80.0, w, h, seed=3)
    mine = sharding.shard_views(n_views, rank, world)
    res = _visible_sets(sc, flat, [views[i] for i in mine], w, h)
    counts = torch.tensor([[c, len(v)] for c, v in res], dtype=torch.int32).reshape(-1, 2)
    ids = torch.from_numpy(np.concatenate([v for _, v in res]) if res else np.zeros(0, np.int32))
    g_counts, g_off, g_ids = sharding.gather_results(counts, ids, n_views, rank, world)
    # the sync-free variant used in the steady state must give the same answer
    raw = sharding.RawGather(n_views, world, 40000, torch.device("cpu"))
    padded = torch.zeros(40000, dtype=torch.int32)
    padded[:ids.shape[0]] = ids
    raw.run(counts, padded)
    assert not raw.overflowed()
    r_counts, r_off, r_ids = raw.reorder()
    assert torch.equal(r_counts, g_counts) and torch.equal(r_off, g_off) and torch.equal(r_ids, g_ids)
    if rank == 0:
        q.put((g_counts.numpy(), g_off.numpy(), g_ids.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [7, 8])
def test_sharded_gather_equals_single_process(n_views):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_views, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, off, ids = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w, h = 128, 80
    sc = helpers.small_scene(n=3000, k=96, extent=80.0, seed=7)
    _, flat = helpers.host_flatten(sc)
    views = scenes.agent_cameras(n_views, 80.0, w, h, seed=3)
    want = _visible_sets(sc, flat, views, w, h)
    assert counts.shape == (n_views, 2)
    for v, (c, vis) in enumerate(want):
        assert counts[v, 0] == c and counts[v, 1] == len(vis)
        assert np.array_equal(ids[off[v]:off[v + 1]], vis)


def test_shard_views_covers_everything_once():
    for n, world in ((4096, 8), (7, 2), (3, 4)):
        got = sorted(sum((sharding.shard_views(n, r, world) for r in range(world)), []))
        assert got == list(range(n))
