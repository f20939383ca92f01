"""Multi-GPU host logic of the visibility path: views shard over ranks, results are gathered.

Views are independent, every rank holds the whole scene, so the data path has no collective.  The only exchange is the gather of
per-view counts and visible sets at the end of a batch, done here with torch.distributed collectives (NCCL over NVLink on GPUs;
the same code runs on gloo/CPU tensors, which is how it is tested without GPUs).
"""
import torch
import torch.distributed as dist


def shard_views(n_views, rank, world):
    """Interleaved assignment (view v -> rank v mod world): neighbouring cameras differ in cost, interleaving balances the ranks."""
    return list(range(rank, n_views, world))


def gather_results(counts_local, ids_local, n_views, rank, world, group=None):
    """All-gather the per-view (query size, visible count) pairs and the packed visible ids of every rank and put them back into global
    view order.  counts_local: int32 [n_local, 2]; ids_local: int32 [sum(counts_local[:,1])], views in local order.
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
    # back to global order: view v lives on rank v % world at local slot v // world
    v = torch.arange(n_views, device=dev)
    r, s = v % world, v // world
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

    def __init__(self, n_views, world, cap, device, group=None):
        self.n_views, self.world, self.cap, self.group = n_views, world, max(int(cap), 1), group
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
        r, s = v % self.world, v // self.world
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
