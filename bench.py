#!/usr/bin/env python
"""bench.py -- views/sec of the per-frame visibility path (occluder raster -> depth hierarchy -> octree query -> visibility).

Workload = BASELINE.json configs[2] ("C3", the configuration the metric is quoted on): 4,096 independent agent cameras over a
1M-drawable / 4,000-occluder scene, 256x256 occlusion buffers, views sharded over the ranks.
A step = one pass of the whole path over the 4,096-view batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--views V]
    torchrun --nproc-per-node N bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value` = views/s with inputs resident in HBM (device time, CUDA events, max over ranks);
`e2e` = the same through one C-ABI call per step with host buffers in and out; `roofline` = the dominant kernel against the
measured HBM peak; `cpu_baseline` = the reference's own CPU code (oracle/_ref) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 256, 256
FLUSH_BYTES = 160 << 20  # > the 126 MB L2
METRIC = "views/sec (1M drawables, 256x256 occlusion)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--views", type=int, default=4096)
    ap.add_argument("--drawables", type=int, default=1_000_000)
    ap.add_argument("--occluders", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush", action="store_true", help="write a 160 MiB buffer before every step even when the working set exceeds L2")
    ap.add_argument("--inflight", type=int, default=0,
                    help="result slots (u3d_batch objects on their own streams) the steps rotate through; 0 = 8 for shards above 1024 views, 12 below "
                         "(a small batch leaves more of the GPU idle at the end of each of its ~25 dependent passes: measured on one GPU with lone "
                         "512-view batches, 8 -> 12 slots: device 425 -> 430 k views/s, end to end 325 -> 355 k, profiles/r01_lone512_slots.jsonl)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 result exchange: p2p = our fused pack+store kernel over NVLink peer memory (u3d_gather_*), nccl = library all-gather")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="views of the workload timed on the host cores")
    ap.add_argument("--shard", default="balanced", choices=["balanced", "interleaved"],
                    help="N>1 view assignment: balanced = equal view counts, per-view cost proxies of one untimed pass equalised over the ranks "
                         "(urho3d_b200.sharding.balance_views); interleaved = view v on rank v mod N")
    return ap.parse_args()


def build_inputs(args):
    from urho3d_b200 import scenes
    extent = scenes.CONFIGS["C3"]["extent"]
    sc = scenes.Scene(args.drawables, args.occluders, extent)
    views = scenes.agent_cameras(args.views, extent, WIDTH, HEIGHT)
    return sc, views


L2_BYTES = 126 << 20


def step_working_set(views_per_rank):
    """Bytes one step of one rank writes and reads besides the replicated scene: depth planes, hierarchies (8 bytes per texel), octant states."""
    return views_per_rank * (4 * WIDTH * HEIGHT + 8 * mip_texels(WIDTH, HEIGHT) + 21856)


def flush_each_step(args, world):
    """Timing rule: flush L2 between timed steps OR use inputs larger than L2.  The per-step working set of every default configuration
    exceeds L2 several times over (and rotates over the batch slots), so no flush is written; a run whose working set is below 1.5 x L2
    (few views) flushes before every step.  --flush forces it."""
    return args.flush or step_working_set((args.views + world - 1) // world) < 1.5 * L2_BYTES


def cache_note(args, world):
    per = (args.views + world - 1) // world
    ws = step_working_set(per)
    if flush_each_step(args, world):
        return "L2 flushed before every timed step (%d MiB write > 126 MB L2, inside the timed region)" % (FLUSH_BYTES >> 20)
    return ("inputs larger than L2, no flush: a step's working set per rank is %.0f MB of depth planes, hierarchies and octant states (%d views) "
            "+ triangle records, queues and result words = %.1f x the 126 MB L2, and consecutive steps run in %d different batch slots; the "
            "replicated scene (36 MB of AABBs + octants), read by every step as in an engine, is meant to stay L2-resident" % (
                ws / 1e6, per, ws / L2_BYTES, max(1, args.inflight)))


def config_dict(args, world):
    return {"workload": "C3: %d agent cameras x %d drawables, %d box occluders (12 tris each), %dx%d occlusion buffers, depth hierarchy, "
                        "occluded octree query + visibility predicate" % (args.views, args.drawables, args.occluders, WIDTH, HEIGHT),
            "views": args.views, "drawables": args.drawables, "occluders": args.occluders, "buffer": [WIDTH, HEIGHT],
            "sharding": ("views interleaved over %d rank(s)" % world if world == 1 or args.shard == "interleaved" else
                         "views cost-balanced over %d ranks: equal view counts, assignment derived from the per-view statistics (submitted / set-up "
                         "triangles, query size, visible count) of one untimed pass, as an engine would use the previous frame's; costly views "
                         "first within a shard" % world),
            "exchange": ("none (1 rank)" if world == 1 else "fused pack + NVLink peer-memory stores + flags (u3d_batch_push_results), every step"
                         if args.gather == "p2p" else "NCCL all_gather_into_tensor of counts and packed ids, every step"),
            "pipelining": "%d batch slots on separate streams; step k runs in slot k %% %d" % (max(1, args.inflight), max(1, args.inflight)),
            "cache": cache_note(args, world)}


# ------------------------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = the UNMODIFIED reference compiled from /root/reference; falls back to the C oracle port)
# ------------------------------------------------------------------------------------------------------------------------------
_HASH_P = np.uint64(0x9E3779B97F4A7C15)


def seq_hash(ids):
    """Order-sensitive 64-bit hash of an id sequence: sum_i (id_i + 1) * P^(i+1) + n * P, in wrapping uint64 arithmetic.  Computed the same
    way from the CPU arm's result_ sequences and from the CUDA path's packed output; swapping two ids or dropping one changes it."""
    ids = np.asarray(ids).astype(np.uint64)
    n = ids.shape[0]
    with np.errstate(over="ignore"):
        w = np.cumprod(np.full(n, _HASH_P, np.uint64), dtype=np.uint64) if n else np.zeros(0, np.uint64)
        return int(((ids + np.uint64(1)) * w).sum(dtype=np.uint64) + np.uint64(n) * _HASH_P)


def _ref_worker(ref, sc, views, conn):
    """Persistent worker: waits for (lo, hi) ranges, runs the reference over those views, answers (seconds, visible, triangles, per-view
    hashes of the visible-id sequences)."""
    while True:
        msg = conn.recv()
        if msg is None:
            os._exit(0)
        lo, hi = msg
        t0 = time.perf_counter()
        nvis = 0
        tris = 0
        seqs = []
        for i in range(lo, hi):
            ref.set_view_raw(views[i])
            vis, counts, _ = ref.run_view(sc)
            nvis += len(vis)
            tris += int(counts[0])
            seqs.append(vis)
        dt = time.perf_counter() - t0  # hashing is the checker's work, not the reference's: outside the timed part
        conn.send((dt, nvis, tris, [seq_hash(x) for x in seqs]))


class CpuReference:
    """Runs the reference's CPU path over a bounded sample of the workload, one process per host core (the views are independent;
    this is the most favourable way to use all host threads -- the reference itself walks its views one after another on the main thread,
    Renderer.cpp:673-686).  Workers are forked once, after the reference scene has been built, and reused for every step."""

    def __init__(self, sc, views, sample):
        import multiprocessing as mp
        from oracle import refbind
        self.sc = sc
        self.views = views[:sample]
        self.cores = max(1, min(os.cpu_count() or 1, 64, len(self.views)))
        self.kind = "reference" if refbind.available() else "port"
        self.workers = []
        if self.kind == "reference":
            ref = refbind.Ref(0)
            ref.octree_set_size(sc.octree_min, sc.octree_max, sc.octree_levels)
            ref.add_drawables(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
            ref.ob_set_size(WIDTH, HEIGHT)
            ctx = mp.get_context("fork")
            for _ in range(self.cores):
                parent, child = ctx.Pipe(True)
                p = ctx.Process(target=_ref_worker, args=(ref, sc, self.views, child), daemon=True)
                p.start()
                child.close()
                self.workers.append((p, parent))

    def close(self):
        for p, conn in self.workers:
            try:
                conn.send(None)
            except OSError:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)
        self.workers = []

    def step(self):
        """One pass over the sample.  Returns (wall seconds, views, visible total, submitted triangles)."""
        n = len(self.views)
        if self.kind == "port":
            return self._step_port()
        t0 = time.perf_counter()
        for w, (_, conn) in enumerate(self.workers):
            conn.send((n * w // self.cores, n * (w + 1) // self.cores))
        nvis = tris = 0
        self.hashes = []
        dt = None
        for _, conn in self.workers:
            _, a, b, hs = conn.recv()
            dt = time.perf_counter() - t0  # wall clock until the last worker's answer arrived (hashing happens after each worker's timed loop)
            nvis += a
            tris += b
            self.hashes += hs
        return dt, n, nvis, tris

    def _step_port(self):
        from oracle import oraclebind
        from urho3d_b200 import capi
        if not hasattr(self, "_port"):
            t = capi.Octree(self.sc.octree_min, self.sc.octree_max, self.sc.octree_levels)
            t.add(self.sc.aabb, self.sc.flags, self.sc.view_mask, self.sc.occludee, self.sc.cast_shadows)
            tree = oraclebind.OracleTree(t.flatten(), self.sc.aabb, self.sc.flags, self.sc.view_mask, self.sc.occludee, self.sc.cast_shadows)
            ob = oraclebind.OracleOB()
            ob.set_size(WIDTH, HEIGHT)
            self._port = (tree, ob)
            self.views = self.views[:256]
            self.cores = 1
        tree, ob = self._port
        t0 = time.perf_counter()
        nvis = tris = 0
        seqs = []
        for v in self.views:
            tris += ob.draw_scene_view(self.sc, v)
            cand = tree.query(1, v.planes, ob, as_ids=False)
            vis = tree.order[tree.check_visibility(ob, cand)]
            nvis += len(vis)
            seqs.append(vis)
        dt = time.perf_counter() - t0
        self.hashes = [seq_hash(x) for x in seqs]
        return dt, len(self.views), nvis, tris

    def describe(self):
        return "%d of the workload's views, split over %d single-threaded processes, each running the %s per view: Clear, AddTriangles for " \
               "every occluder passing IsInsideFast, DrawTriangles, BuildDepthHierarchy, OccludedFrustumOctreeQuery, IsVisible predicate" % (
                   len(self.views), self.cores, "compiled reference classes (oracle/_ref)" if self.kind == "reference" else "C oracle port")


def _engine_threads_child(conn, sc, views, threads):
    from oracle import refbind
    ref = refbind.Ref(threads)
    ref.octree_set_size(sc.octree_min, sc.octree_max, sc.octree_levels)
    ref.add_drawables(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    ref.ob_set_size(WIDTH, HEIGHT, threaded=True)
    best = None
    for _ in range(2):  # first pass warms caches and the worker threads
        t0 = time.perf_counter()
        for v in views:
            ref.set_view_raw(v)
            ref.run_view(sc)
        best = time.perf_counter() - t0
    conn.send((best, ref.num_threads()))
    conn.close()
    os._exit(0)


def engine_threaded_rate(sc, views, sample=128):
    """The reference the way the engine itself runs it: one process, views one after another (Renderer.cpp:673-686), each view's
    occluder batches rasterised into per-thread buffers and its visibility tests split over WorkQueue worker threads (cores - 1 of them, as
    Engine.cpp:210 creates).  A forked child, so that the worker threads never live in this process.  Returns a dict or None."""
    import multiprocessing as mp
    from oracle import refbind
    if not refbind.available():
        return None
    threads = max(1, min(os.cpu_count() or 1, 64) - 1)
    ctx = mp.get_context("fork")
    parent, child = ctx.Pipe(False)
    p = ctx.Process(target=_engine_threads_child, args=(child, sc, views[:sample], threads), daemon=True)
    p.start()
    child.close()
    try:
        got = parent.recv() if parent.poll(60) else None
    except EOFError:
        got = None
    if p.is_alive():
        p.kill()
    p.join(10)
    if not got:
        return None
    return {"value": len(views[:sample]) / got[0], "unit": "views/s", "worker_threads": got[1], "views": len(views[:sample]),
            "how": "one process, views in sequence, threaded OcclusionBuffer + WorkQueue split of the visibility tests (the engine's own mode)"}


def host_description():
    """nproc and CPU model of the box the CPU arm runs on (SURVEY 8d asks for both next to the reference's number)."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    sc, views = build_inputs(args)
    cpu = CpuReference(sc, views, min(args.cpu_sample, args.views))
    for _ in range(max(args.warmup, 1)):
        cpu.step()
    total = 0.0
    nviews = 0
    tris = 0
    for _ in range(args.steps):
        dt, n, _, t = cpu.step()
        total += dt
        nviews += n
        tris += t
    value = nviews / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32",
            "data": "synthetic", "config": config_dict(args, world), "occluder_mtri_per_s": tris / total / 1e6,
            "cpu_baseline": dict({"value": value, "unit": "views/s", "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.describe()}, **host_description()),
            "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    cpu.close()
    eng = engine_threaded_rate(sc, views)
    if eng:
        line["cpu_baseline"]["engine_threads"] = eng
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------------------
def mip_texels(w, h):
    tot = 0
    while True:
        w, h = (w + 1) // 2, (h + 1) // 2
        tot += w * h
        if w <= 8 and h <= 8:
            return tot


class _DevArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<i4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def balanced_shards(batch, stream, my_ids, n_views, rank, world, device):
    """One untimed pass over this rank's (interleaved) views, per-view statistics exchanged, cost-balanced assignment derived (the same on
    every rank).  Returns (assignment, summary of the proxies' max-over-mean before and after)."""
    import torch
    from urho3d_b200 import capi, sharding
    batch.run(capi.STAGE_VISIBLE, stream)
    n = len(my_ids)
    c_probe = batch.counts(stream).astype(np.int64)[:n]
    t_probe = batch.tri_stats(stream).astype(np.int64)[:n]
    local = torch.from_numpy(np.ascontiguousarray(np.concatenate([t_probe[:, [0, 2]], c_probe], axis=1))).to(device)
    view_stats = sharding.gather_view_stats(local, my_ids, n_views, world).cpu().numpy()
    assignment = sharding.balance_views(view_stats, world)
    # within a shard: costly views first.  CTAs are handed out in view order, and with only a wave or two of CTAs per launch a costly view that
    # starts last is what the launch (and every launch queued behind it on the stream) waits for.
    weight = (view_stats / np.maximum(view_stats.sum(axis=0), 1)).sum(axis=1)
    assignment = [a[np.lexsort((a, -weight[a]))] for a in assignment]

    def worst(parts):
        return [round(float(max(view_stats[p, c].sum() for p in parts) / max(view_stats[:, c].sum() / world, 1.0)), 4)
                for c in range(view_stats.shape[1])]

    summary = {"proxies": ["submitted_tris", "setup_tris", "query_result", "visible"],
               "max_over_mean_interleaved": worst([np.arange(r, n_views, world) for r in range(world)]),
               "max_over_mean_balanced": worst(assignment)}
    return assignment, summary


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.inflight <= 0:
        args.inflight = 8 if (args.views + world - 1) // world > 1024 else 12
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from urho3d_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    args.warmup = max(args.warmup, 3)

    sc, views = build_inputs(args)
    my_ids = list(range(rank, args.views, world))
    my_views = [views[i] for i in my_ids]
    V = len(my_views)

    ctx = capi.Context(local_rank)
    for kv in filter(None, os.environ.get("U3D_DEBUG_OPTS", "").split(",")):  # tuning sweeps only, e.g. "heavy_extent=64"
        name, val = kv.split("=")
        capi.debug_set_option(name, int(val))
    tree = capi.Octree(sc.octree_min, sc.octree_max, sc.octree_levels)
    tree.add(sc.aabb, sc.flags, sc.view_mask, sc.occludee, sc.cast_shadows)
    gscene = capi.GpuScene(ctx, octree=tree)
    mesh = capi.Mesh(ctx, sc.mesh_vtx, sc.mesh_stride, sc.mesh_idx)
    occ = capi.Occluders(ctx, sc.occ_aabb, sc.occ_model, [mesh], cull_mode=sc.cull_mode)
    out_cap = 32768
    batch = capi.Batch(ctx, gscene, occ, WIDTH, HEIGHT, V, out_cap, tri_pool=V * 14000)
    packed = capi.pack_views(my_views)
    # a non-default torch stream: its handle is passed to every C-ABI call, and torch.cuda.Event records on the same stream
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    do_flush = flush_each_step(args, world)

    # N>1: cost-balanced shards.  Views differ widely in cost (std ~ mean), so 512 interleaved views per rank leave 5-7 % between the ranks of an
    # 8-GPU run and every step ends with the slowest one.  One untimed pass on the interleaved shards yields per-view statistics; they are
    # exchanged and every rank derives the same balanced assignment (equal view counts, so no buffer or exchange layout changes).
    assignment, shard_balance = None, None
    if world > 1 and args.shard == "balanced":
        batch.set_views(packed, stream)
        assignment, shard_balance = balanced_shards(batch, stream, my_ids, args.views, rank, world, torch.device("cuda", local_rank))
        my_ids = [int(i) for i in assignment[rank]]
        assert len(my_ids) == V, "the balanced assignment keeps every rank's view count"
        my_views = [views[i] for i in my_ids]
        packed = capi.pack_views(my_views)

    # device-resident leg ----------------------------------------------------------------------------------------------------
    # `inflight` result slots, each a u3d_batch on its own stream: step k runs in slot k % inflight, so the tail of one step's kernels (a few
    # CTAs left on a mostly idle GPU) and, for N>1, its NCCL gather overlap the next step's kernels.  All K steps start after the start event
    # and finish before the end event.
    inflight = max(1, args.inflight)
    do_gather = world > 1 and not os.environ.get("U3D_BENCH_NO_GATHER")  # diagnostic only: a run without the gather is not a bench value
    batch.set_views(packed, stream)
    slots = [dict(batch=batch, tstream=tstream)]
    for _ in range(inflight - 1):
        st_i = torch.cuda.Stream()
        b_i = capi.Batch(ctx, gscene, occ, WIDTH, HEIGHT, V, out_cap, tri_pool=V * 14000)
        b_i.set_views(packed, st_i.cuda_stream)
        slots.append(dict(batch=b_i, tstream=st_i))
    for sl in slots:
        sl["stream"] = sl["tstream"].cuda_stream
        sl["done"] = None
        sl["batch"].set_pipelined(inflight > 1)
    comm = None
    if world > 1:
        from urho3d_b200 import sharding
        # The gather buffers are sized from one untimed pass, so no host synchronisation is needed inside the timed steps.
        batch.run(capi.STAGE_VISIBLE, stream)
        c0 = batch.counts(stream)
        cap_t = torch.tensor([int(c0[:, 1].sum() * 1.25) + 4096], dtype=torch.int64, device="cuda")
        dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
        cap = min(int(cap_t.item()), V * out_cap)
        comm = torch.cuda.Stream()  # one stream for every slot's collectives: all ranks issue them in the same order
        per = (args.views + world - 1) // world
        if args.gather == "p2p":
            # receive block of every slot on every rank; the 64-byte IPC handles are exchanged once, outside the timed region.  If peer memory
            # cannot be mapped on this box (no P2P between the devices, IPC disabled in the container) every rank falls back to the NCCL gather.
            def agreed(flag):
                t = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                return bool(t.item())

            exported = []
            try:  # local: allocate and export
                for sl in slots:
                    g = capi.Gather(ctx, world, rank, per, cap)
                    exported.append((g, g.export()))
                ok = True
            except capi.U3DError as e:
                sys.stderr.write("rank %d: peer-memory exchange unavailable (%s)\n" % (rank, e))
                ok = False
            if agreed(ok):
                all_handles = [None] * world
                dist.all_gather_object(all_handles, [h for _, h in exported])  # collective: every rank takes part
                try:  # local: map the peers' blocks
                    for i, (g, _) in enumerate(exported):
                        g.connect([all_handles[r][i] for r in range(world)])
                        slots[i]["gather"] = g
                except capi.U3DError as e:
                    sys.stderr.write("rank %d: peer-memory exchange unavailable (%s)\n" % (rank, e))
                    ok = False
                ok = agreed(ok)
            else:
                ok = False
            if not ok:
                args.gather = "nccl"
        for sl in slots:
            if args.gather == "p2p":
                continue
            cptr, _, _ = sl["batch"].device_results()
            sl["counts"] = torch.as_tensor(_DevArray(cptr, (V, 2)), device="cuda")
            sl["batch"].run(capi.STAGE_VISIBLE, sl["stream"])
            _, pptr = sl["batch"].pack_device(sl["stream"])  # allocates the packed buffer once; its address is stable afterwards
            sl["packed"] = torch.as_tensor(_DevArray(pptr, (V * out_cap,)), device="cuda")
            sl["gather"] = sharding.RawGather(args.views, world, cap, torch.device("cuda", local_rank), assignment=assignment)
    torch.cuda.synchronize()

    def step_device(k):
        sl = slots[k % len(slots)]
        ts = sl["tstream"]
        with torch.cuda.stream(ts):
            if sl["done"] is not None:
                ts.wait_event(sl["done"])  # this slot's previous gather has read its buffers
            if do_flush:
                flush.zero_()
            sl["batch"].run(capi.STAGE_VISIBLE, sl["stream"])
            extra = 0
            if do_gather and args.gather == "p2p":
                # the only exchange of the path: this rank's counts and packed visible sets are stored into every rank's receive
                # block by our own kernel (pack + NVLink peer stores + arrival flag); the comm stream waits for all ranks' flags and releases
                g = sl["gather"]
                g.push(sl["batch"], sl["stream"])
                pushed = torch.cuda.Event()
                pushed.record(ts)
                comm.wait_event(pushed)  # program order only: the flags carry the real dependency
                g.wait(comm.cuda_stream)
                g.release(comm.cuda_stream)
                sl["done"] = torch.cuda.Event()
                sl["done"].record(comm)
                extra = 3  # k_push_results, k_gather_wait, k_gather_signal
            elif do_gather:
                # library baseline: every rank's per-view counts and packed visible sets, all-gathered over NCCL
                sl["batch"].pack_device(sl["stream"])
                ready = torch.cuda.Event()
                ready.record(ts)
                comm.wait_event(ready)
                with torch.cuda.stream(comm):
                    sl["gather"].run(sl["counts"], sl["packed"])
                    sl["done"] = torch.cuda.Event()
                    sl["done"].record(comm)
                extra = 2
        return sl["batch"].last_launches() + extra

    def fork(ev):
        for sl in slots[1:]:
            sl["tstream"].wait_event(ev)

    def join():
        # the main stream waits for everything the slots (and the gathers) were given
        for sl in slots:
            if sl["done"] is not None:
                tstream.wait_event(sl["done"])
            if sl["tstream"] is not tstream:
                e = torch.cuda.Event()
                e.record(sl["tstream"])
                tstream.wait_event(e)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_clock0 = time.perf_counter()
    for k in range(max(args.warmup, len(slots))):
        step_device(k)
    join()
    torch.cuda.synchronize()
    counts = batch.counts(stream)  # also checks the overflow flags
    stats = batch.tri_stats(stream)
    for sl in slots[1:]:
        assert np.array_equal(sl["batch"].counts(sl["stream"]), counts), "slots disagree"
    if do_gather and args.gather == "p2p":
        all_counts = [torch.zeros((V, 2), dtype=torch.int64, device="cuda") for _ in range(world)] if V * world == args.views else None
        if all_counts is not None:
            dist.all_gather(all_counts, torch.from_numpy(counts.astype(np.int64)).cuda())
        for sl in slots:
            g_counts, g_offsets, g_ids = sl["gather"].read(comm.cuda_stream)  # raises on overflow / timeout
            assert np.array_equal(g_counts[rank, :V], counts), "exchanged counts disagree with the local ones"
            if all_counts is not None:  # every other rank's counts arrived as that rank computed them
                for r in range(world):
                    assert np.array_equal(g_counts[r, :V], all_counts[r].cpu().numpy().astype(np.uint32)), "rank %d counts differ" % r
            off, ids_local = sl["batch"].packed(sl["stream"], capacity=int(counts[:, 1].sum()))
            assert np.array_equal(g_offsets[rank, :V + 1], off) and np.array_equal(g_ids[rank, :off[-1]], ids_local), "exchanged ids differ"
    elif do_gather:
        for sl in slots:
            assert not sl["gather"].overflowed(), "gather buffers too small"
            g_counts, _, _ = sl["gather"].reorder()
            assert np.array_equal(g_counts[my_ids].cpu().numpy().astype(np.uint32), counts), "gathered counts disagree with the local ones"

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    # One event pair around the K steps.  (If the working set were below 1.5 x L2 a 160 MiB flush write would run before each step, inside the
    # timed region: flush_each_step.)
    ev_start, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ev_start.record(tstream)
    fork(ev_start)
    for k in range(args.steps):
        launches += step_device(k)
    t_enqueued = time.perf_counter()  # host time to enqueue the K steps: the step is launch-bound when this approaches the device time
    join()
    ev_end.record(tstream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_end = time.perf_counter()
    total_ms = torch.tensor([ev_start.elapsed_time(ev_end)], dtype=torch.float64, device="cuda")
    rank_ms = None
    if world > 1:
        every = [torch.zeros_like(total_ms) for _ in range(world)]
        dist.all_gather(every, total_ms)
        rank_ms = [round(float(t.item()) / args.steps, 4) for t in every]
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = None

    # per-phase device times (events after each kernel phase), same inputs
    phase_ms = {}
    for _ in range(3):
        flush.zero_()
        torch.cuda.synchronize()
        for name, ms in batch.run_timed(capi.STAGE_VISIBLE, stream).items():
            phase_ms.setdefault(name, []).append(ms)
    phase_ms = {k: statistics.median(v) for k, v in phase_ms.items()}

    # e2e leg: one C-ABI call per step, host buffers in (views) and out (counts + packed visible ids).  One host thread per slot, as an engine
    # with worker threads would submit batches: each call is synchronous (returns with the results in host memory), calls of different slots
    # overlap on the device.  Wall clock over all K calls. ------------------------------------------------------------------------------
    import threading
    n_ids = int(counts[:, 1].sum() * 1.25) + 1024
    for sl in slots:
        sl["h_counts"] = torch.empty((V, 2), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        sl["h_offsets"] = torch.empty(V + 1, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        sl["h_ids"] = torch.empty(n_ids, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        h_views = torch.empty(packed.nbytes, dtype=torch.uint8).pin_memory().numpy()
        h_views[:] = packed.view(np.uint8).reshape(-1)
        sl["h_packed"] = h_views.view(capi.VIEW_DTYPE)

    def e2e_call(sl):
        sl["batch"].cull_views(sl["h_packed"], capi.STAGE_VISIBLE, sl["stream"], sl["h_counts"], sl["h_offsets"], sl["h_ids"])

    def e2e_worker(t, steps, errs, gate):
        try:
            torch.cuda.set_device(local_rank)
            gate.wait()  # the engine's worker threads exist before the frame starts: their creation is not part of a step
            for k in range(t, steps, len(slots)):
                e2e_call(slots[t])
        except Exception as e:  # noqa: BLE001 - re-raised on the main thread
            errs.append(e)
            gate.abort()

    def e2e_run(steps):
        errs = []
        gate = threading.Barrier(len(slots) + 1)
        threads = [threading.Thread(target=e2e_worker, args=(t, steps, errs, gate)) for t in range(len(slots))]
        for th in threads:
            th.start()
        try:
            gate.wait()
        except threading.BrokenBarrierError:
            pass
        t0 = time.perf_counter()
        for th in threads:
            th.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if errs:
            raise errs[0]
        return dt

    # K steps shared by one thread per slot is one or two calls per thread at small K: the e2e leg runs at least 8 calls per slot (its own
    # step count is reported as e2e.steps) so that it measures the steady state of the call path, like `value` does for the kernels
    e2e_steps = max(args.steps, 8 * len(slots))
    if world * len(slots) > (os.cpu_count() or 1):
        # more calling threads on the box than host cores: wait on a blocking event instead of spinning (what an engine with that many
        # worker threads would configure)
        capi.debug_set_option("blocking_wait", 1)
        e2e_wait = "blocking event (%d calling threads on %d host cores)" % (world * len(slots), os.cpu_count() or 1)
    else:
        e2e_wait = "cudaStreamSynchronize"
    e2e_run(2 * len(slots))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([e2e_run(e2e_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    h_counts, h_offsets = slots[0]["h_counts"], slots[0]["h_offsets"]
    for sl in slots:
        assert np.array_equal(sl["h_counts"], counts), "e2e leg disagrees with the device-resident leg"
    if sampler:
        # keep the same load on for >= 1.5 s in total so that nvidia-smi (100 ms period) gets enough samples under load
        while time.perf_counter() - t_clock0 < 1.5:
            batch.run(capi.STAGE_VISIBLE, stream)  # rank-local work only: no collective here (the other ranks do not run this loop)
            torch.cuda.synchronize()
        # nvidia-smi samples every 100 ms: the window spans warm-up, the timed steps, the phase timing and the e2e leg (GPU busy throughout)
        clocks = sampler.stop(t_clock0, time.perf_counter())
        clocks["window"] = "warm-up + timed steps + e2e leg"
    h2d = packed.nbytes
    d2h = h_counts.nbytes + h_offsets.nbytes + int(h_offsets[-1]) * 4
    tot = torch.tensor([h2d, d2h, int(stats[:, 0].sum()), int(counts[:, 1].sum()), int(counts[:, 0].sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    h2d, d2h, sub_tris, n_vis, n_cand = [float(x) for x in tot.tolist()]

    if rank == 0:
        value = args.views * args.steps / (total_ms * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        N, M = gscene.num_drawables, gscene.num_octants
        mips = mip_texels(WIDTH, HEIGHT)
        vis_per_view = n_vis / args.views
        tri_per_view = sub_tris / args.views
        # algorithmic bytes per view, SURVEY 8(d)
        b_cull = 32.0 * N + 32.0 * M + 4.0 * vis_per_view + 4.0 + 4.0 * WIDTH * HEIGHT + 8.0 * mips
        b_tri = 3 * 2 + 36 + 48.0 / 12.0
        cull_keys = ("tree", "frustum", "occlusion", "emit")
        raster_keys = ("clear", "select", "scan", "setup", "fill", "second_pass")
        cull_ms = sum(phase_ms.get(k, 0.0) for k in cull_keys)
        raster_ms = sum(phase_ms.get(k, 0.0) for k in raster_keys)
        phases = {k: {"ms": round(v, 4)} for k, v in phase_ms.items()}
        phases["cull_total"] = {"ms": round(cull_ms, 4), "alg_GBps": round(b_cull * V / max(cull_ms, 1e-6) / 1e6, 1)}
        phases["raster_total"] = {"ms": round(raster_ms, 4),
                                  "alg_GBps": round((tri_per_view * b_tri + 4.0 * WIDTH * HEIGHT + 8.0 * mips) * V / max(raster_ms, 1e-6) / 1e6, 1)}
        achieved = b_cull * V / (cull_ms * 1e-3) / 1e9
        # Work-based bounds (the flat-scan fraction is a convention, not a bound): what the kernels actually execute and touch,
        # per view, from the committed ncu summaries (profiles/r02_work.json; counters of one 4096-view launch divided by its views), against
        # the lane-instruction rate of the chip (SMs x 128 lanes x SM clock) and the L2 bandwidth.
        work = None
        work_path = os.path.join(ROOT, "profiles", "r02_work.json")
        if os.path.exists(work_path):
            wj = json.load(open(work_path))
            sm_clock_hz = 1e6 * float((clocks or {}).get("sm_mhz") or wj.get("sm_mhz", 1965.0))
            lane_rate = wj.get("sms", 148) * 128 * sm_clock_hz
            l2_peak = wj.get("l2_bytes_per_clk", 6300.0) * sm_clock_hz
            work = {"source": "profiles/r02_work.json (ncu --set full, one launch per kernel)", "lane_inst_per_s_peak": lane_rate,
                    "l2_GBps_peak": l2_peak / 1e9, "kernels": {}}
            for kname, kw in wj.get("kernels", {}).items():
                ms = phase_ms.get(kw.get("phase", ""), 0.0) * kw.get("phase_share", 1.0)
                if ms <= 0.0:
                    continue
                ent = {"phase": kw.get("phase"), "ms": round(ms, 4),
                       "lane_inst_per_view": kw["thread_inst_per_view"],
                       "issue_frac": kw["thread_inst_per_view"] * V / (ms * 1e-3) / lane_rate,
                       "l2_bytes_per_view": kw["lts_bytes_per_view"],
                       "l2_frac": kw["lts_bytes_per_view"] * V / (ms * 1e-3) / l2_peak,
                       "dram_bytes_per_view": kw.get("dram_bytes_per_view")}
                if "smem_atomics_per_view" in kw:
                    ent["smem_atomics_per_view"] = kw["smem_atomics_per_view"]
                    # shared-memory wavefronts of atomic instructions against one wavefront per clock and SM
                    ent["smem_atomic_frac"] = kw["smem_atomics_per_view"] * V / (ms * 1e-3) / (wj.get("sms", 148) * sm_clock_hz)
                    if "smem_wavefronts_per_view" in kw:
                        ent["smem_frac"] = kw["smem_wavefronts_per_view"] * V / (ms * 1e-3) / (wj.get("sms", 148) * sm_clock_hz)
                work["kernels"][kname] = ent
        dram_per_view = None
        if work:
            dram_per_view = sum((k.get("dram_bytes_per_view") or 0.0) for k in work["kernels"].values() if k["phase"] in cull_keys) or None
        line = {"metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32",
                "data": "synthetic", "config": config_dict(args, world),
                "occluder_mtri_per_s": sub_tris * args.steps / (total_ms * 1e-3) / 1e6,
                # SURVEY 8d's definition: submitted triangles / time of the draw + hierarchy phases alone (rank 0's shard, phase timers)
                "occluder_mtri_per_s_raster_phases": float(stats[:, 0].sum()) / max(1e-9, 1e-3 * raster_ms) / 1e6,
                "per_view": {"submitted_tris": tri_per_view, "query_result": n_cand / args.views, "visible": vis_per_view, "octants": M},
                "e2e": {"value": args.views * e2e_steps / e2e_s, "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "host_wait": e2e_wait},
                "gpu_launches": launches, "host_enqueue_ms_per_step": round((t_enqueued - t_begin) * 1e3 / args.steps, 4),
                "roofline": {"bound": "hbm", "kernel": "cull = k_dense_tree + k_dense_frustum + k_dense_occlusion_walk + k_dense_emit (phases tree, frustum, "
                                                       "occlusion, emit)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (dram_per_view * V if dram_per_view else None),
                             "peak_source": peak_src, "algorithmic_bytes_per_view": b_cull, "views_per_launch": V,
                             "traffic_source": "profiles/r02_work.json: dram__bytes_read.sum + dram__bytes_write.sum of the cull kernels' ncu captures, per view",
                             "note": "algorithmic bytes follow SURVEY 8(d)'s flat-scan CONVENTION (every AABB read once per view); the octree walk "
                                     "skips ~99% of them and the scene SoA is L2-resident, so frac > 1 is expected and is not a bandwidth claim. "
                                     "The bounds that say how far the kernels are from the machine are in `work`: executed lane instructions against "
                                     "SMs x 128 x clock, L2 bytes against the L2 bandwidth, shared-memory atomics for the fill kernel",
                             "work": work},
                "phases_rank0": phases, "clocks": clocks}
        if rank_ms is not None:
            line["ms_per_step_by_rank"] = rank_ms  # the exchange couples the ranks every step, so these differ little; `value` uses the max
        if shard_balance is not None:
            line["shard_balance"] = shard_balance
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuReference(sc, views, min(args.cpu_sample, args.views))
            cpu.step()  # warm-up pass (page faults of the forked workers, caches), as the --impl reference arm does
            dt, n, nvis_cpu, tris_cpu = cpu.step()
            line["cpu_baseline"] = dict({"value": n / dt, "unit": "views/s", "cores": cpu.cores, "kind": cpu.kind, "sample": cpu.describe()},
                                        **host_description())
            # parity gate on the real workload: every view of the sample through the CPU arm and through the CUDA path (the e2e leg's
            # host output) must give the same visible-id SEQUENCE (order-sensitive 64-bit hash per view, Octree.cpp:243-254's result_ order), and
            # the same totals of visible drawables and submitted occluder triangles
            n_cmp = len(cpu.views)
            h_ids = slots[0]["h_ids"]
            gpu_hashes = [seq_hash(h_ids[int(h_offsets[i]):int(h_offsets[i + 1])]) for i in range(n_cmp)]
            n_equal = sum(1 for a, b in zip(cpu.hashes, gpu_hashes) if a == b)
            line["parity_check"] = {"views": n_cmp, "against": cpu.kind,
                                    "views_equal": "%d/%d" % (n_equal, n_cmp),
                                    "how": "order-sensitive 64-bit hash of each view's visible drawable id sequence, CPU arm vs the e2e leg's host output",
                                    "visible": [int(nvis_cpu), int(counts[:n_cmp, 1].sum())],
                                    "submitted_tris": [int(tris_cpu), int(stats[:n_cmp, 0].sum())],
                                    "equal": bool(n_equal == n_cmp and len(cpu.hashes) == n_cmp and int(nvis_cpu) == int(counts[:n_cmp, 1].sum())
                                                  and int(tris_cpu) == int(stats[:n_cmp, 0].sum()))}
            cpu.close()
            eng = engine_threaded_rate(sc, views)
            if eng:
                line["cpu_baseline"]["engine_threads"] = eng
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
