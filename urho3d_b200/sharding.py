"""Multi-GPU host logic of the visibility path: views shard over ranks, results are gathered.

Views are independent, every rank holds the whole scene, so the data path has no collective.  The only exchange is the gather of
per-view counts and visible sets at the end of a batch, done here with torch.distributed collectives (NCCL over NVLink on GPUs;
the same code runs on gloo/CPU tensors, which is how it is tested without GPUs).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_views(n_views, rank, world):
    """Interleaved assignment (view v -> rank v mod world): neighbouring cameras differ in cost, interleaving balances the ranks."""
    return list(range(rank, n_views, world))


def balance_views(costs, world):
    """Cost-balanced assignment.  Views differ widely in cost (submitted occluder triangles, query survivors, visible drawables: the standard
    deviation is about the mean on the benchmark scene), so with a few hundred views per rank the interleaved shards differ by 5-7 % at 8
    ranks, and a batch ends with its slowest rank.  `costs` is [n_views] or [n_views, D]: per-view cost proxies, normally the previous
    frame's per-view statistics (cameras move little between frames).  Every column is balanced at once: views are taken in decreasing order
    of their normalised total and given to the rank whose worst column stays lowest, while every rank keeps exactly the number of views
    the interleaved rule would give it (so buffer sizes and the exchange layout do not change); a few swaps polish the result.  Deterministic: every rank computes the same
    answer from the same statistics.  Returns one ascending int64 array of view indices per rank."""
    c = np.asarray(costs, dtype=np.float64)
    if c.ndim == 1:
        c = c[:, None]
    n = c.shape[0]
    c = np.where(np.isfinite(c) & (c > 0), c, 0.0)
    tot = c.sum(axis=0)
    c = c[:, tot > 0] / tot[tot > 0]
    if c.shape[1] == 0:                                        # no information: the interleaved rule
        return [np.arange(r, n, world, dtype=np.int64) for r in range(world)]
    key = c.sum(axis=1)
    order = np.lexsort((np.arange(n), -key))                   # decreasing cost, ties by view index
    quota = np.array([len(range(r, n, world)) for r in range(world)])
    load = np.zeros((world, c.shape[1]))
    held = np.zeros(world, dtype=np.int64)
    owner = np.empty(n, dtype=np.int64)
    for v in order:
        worst = (load + c[v]).max(axis=1)
        worst[held >= quota] = np.inf
        r = int(np.argmin(worst))                              # first minimum: ties go to the lower rank
        owner[v] = r
        load[r] += c[v]
        held[r] += 1
    # The quotas force the last (cheap) views onto whichever ranks still have room; a few swaps between the rank holding the worst column
    # and the rank lightest in that column take out what that leaves.
    for _ in range(4 * world):
        a, d = np.unravel_index(int(np.argmax(load)), load.shape)
        b = int(np.argmin(load[:, d]))
        if a == b:
            break
        ia, ib = np.flatnonzero(owner == a), np.flatnonzero(owner == b)
        if not len(ia) or not len(ib):
            break
        ia, ib = ia[::max(1, len(ia) // 512)], ib[::max(1, len(ib) // 512)]  # a bounded number of candidate pairs is plenty
        delta = c[ib][None, :, :] - c[ia][:, None, :]            # [|a|, |b|, D]: what rank a gains when it trades view i for view j
        after = np.maximum((load[a] + delta).max(axis=2), (load[b] - delta).max(axis=2))
        i, j = np.unravel_index(int(np.argmin(after)), after.shape)
        if after[i, j] >= max(load[a].max(), load[b].max()) * (1.0 - 1e-12):
            break
        load[a] += delta[i, j]
        load[b] -= delta[i, j]
        owner[ia[i]], owner[ib[j]] = b, a
    return [np.flatnonzero(owner == r).astype(np.int64) for r in range(world)]


def gather_view_stats(stats_local, my_views, n_views, world, group=None):
    """All-gather per-view statistics (int64 [n_local, D], rows in the order of `my_views`) into global view order: [n_views, D] on every
    rank.  `my_views` is this rank's assignment; the views' indices travel with the rows, so any assignment works."""
    dev = stats_local.device
    per = (n_views + world - 1) // world
    d = stats_local.shape[1]
    pad = torch.full((per, d + 1), -1, dtype=torch.int64, device=dev)
    n_local = stats_local.shape[0]
    pad[:n_local, 0] = torch.as_tensor(np.asarray(my_views, dtype=np.int64), device=dev)
    pad[:n_local, 1:] = stats_local.to(torch.int64)
    everything = torch.empty((world * per, d + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(everything, pad, group=group)
    everything = everything[everything[:, 0] >= 0]
    out = torch.zeros((n_views, d), dtype=torch.int64, device=dev)
    out[everything[:, 0]] = everything[:, 1:]
    return out


def _placement(n_views, world, assignment, device):
    """(owner rank, local slot) of every view.  Default = the interleaved rule."""
    if assignment is None:
        v = torch.arange(n_views, device=device)
        return v % world, v // world
    r = torch.empty(n_views, dtype=torch.int64)
    s = torch.empty(n_views, dtype=torch.int64)
    seen = 0
    for rank, ids in enumerate(assignment):
        ids = torch.as_tensor(np.asarray(ids, dtype=np.int64))
        r[ids] = rank
        s[ids] = torch.arange(ids.shape[0])
        seen += ids.shape[0]
    assert seen == n_views, "the assignment must cover every view exactly once"
    return r.to(device), s.to(device)


def gather_results(counts_local, ids_local, n_views, rank, world, group=None, assignment=None):
    """All-gather the per-view (query size, visible count) pairs and the packed visible ids of every rank and put them back into global
    view order.  counts_local: int32 [n_local, 2]; ids_local: int32 [sum(counts_local[:,1])], views in local order.
    `assignment` (one index list per rank, e.g. from balance_views) replaces the interleaved rule.
    Returns (counts [n_views, 2], offsets [n_views + 1], ids) as tensors on the input device, identical on every rank."""
    dev = counts_local.device
    per = (n_views + world - 1) // world
    n_local = counts_local.shape[0]
    pad = torch.zeros((per, 2), dtype=torch.int32, device=dev)
    pad[:n_local] = counts_local
    all_counts = torch.empty((world * per, 2), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_counts, pad, group=group)
    all_counts = all_counts.view(world, per, 2)
    totals = all_counts[:, :, 1].sum(dim=1)                    # ids held by each rank
    cap = int(totals.max().item())                             # one host sync per batch
    send = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    send[:ids_local.shape[0]] = ids_local
    all_ids = torch.empty(world * max(cap, 1), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(all_ids, send, group=group)
    all_ids = all_ids.view(world, max(cap, 1))
    # back to global order: view v lives on rank v % world at local slot v // world (or where the assignment put it)
    v = torch.arange(n_views, device=dev)
    r, s = _placement(n_views, world, assignment, dev)
    counts = all_counts[r, s]
    local_off = torch.cumsum(all_counts[:, :, 1], dim=1) - all_counts[:, :, 1]   # exclusive, per rank
    src_off = local_off[r, s].to(torch.int64)
    vis = counts[:, 1].to(torch.int64)
    offsets = torch.zeros(n_views + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(vis, dim=0)
    total = int(offsets[-1].item())
    if total == 0:
        return counts, offsets, torch.zeros(0, dtype=torch.int32, device=dev)
    view_of = torch.repeat_interleave(v, vis)
    within = torch.arange(total, device=dev) - offsets[view_of]
    ids = all_ids[r[view_of], src_off[view_of] + within]
    return counts, offsets, ids


class RawGather:
    """Steady-state gather with no host synchronisation: two all_gather_into_tensor calls on fixed-size buffers (per-rank padded counts and
    per-rank packed ids up to `cap`).  Every rank ends up with every rank's packed lists; view v is found at rank v % world, local slot
    v // world.  `reorder()` turns that into global view order when a consumer wants it (it synchronises)."""

    def __init__(self, n_views, world, cap, device, group=None, assignment=None):
        self.n_views, self.world, self.cap, self.group = n_views, world, max(int(cap), 1), group
        self.assignment = assignment  # None = interleaved; else one view-index list per rank (balance_views)
        self.per = (n_views + world - 1) // world
        self.pad_counts = torch.zeros((self.per, 2), dtype=torch.int32, device=device)
        self.all_counts = torch.empty((world * self.per, 2), dtype=torch.int32, device=device)
        self.all_ids = torch.empty(world * self.cap, dtype=torch.int32, device=device)

    def run(self, counts_local, packed_local):
        """counts_local: int32 [n_local, 2]; packed_local: int32 tensor with at least `cap` elements (only the first sum(counts) are meaningful)."""
        self.pad_counts[:counts_local.shape[0]].copy_(counts_local)
        dist.all_gather_into_tensor(self.all_counts, self.pad_counts, group=self.group)
        dist.all_gather_into_tensor(self.all_ids, packed_local[:self.cap], group=self.group)

    def overflowed(self):
        totals = self.all_counts.view(self.world, self.per, 2)[:, :, 1].sum(dim=1)
        return bool((totals > self.cap).any().item())

    def reorder(self):
        dev = self.all_counts.device
        ac = self.all_counts.view(self.world, self.per, 2)
        ids2 = self.all_ids.view(self.world, self.cap)
        v = torch.arange(self.n_views, device=dev)
        r, s = _placement(self.n_views, self.world, self.assignment, dev)
        counts = ac[r, s]
        local_off = torch.cumsum(ac[:, :, 1], dim=1) - ac[:, :, 1]
        vis = counts[:, 1].to(torch.int64)
        offsets = torch.zeros(self.n_views + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(vis, dim=0)
        total = int(offsets[-1].item())
        if total == 0:
            return counts, offsets, torch.zeros(0, dtype=torch.int32, device=dev)
        view_of = torch.repeat_interleave(v, vis)
        within = torch.arange(total, device=dev) - offsets[view_of]
        return counts, offsets, ids2[r[view_of], local_off[r, s].to(torch.int64)[view_of] + within]
