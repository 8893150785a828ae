"""Multi-GPU host logic: how edge batches and scenarios are split over ranks and how the per-rank
prune records come back.  One process per GPU; edges are independent given
the replicated read-only world state, so there is NO data-path collective -- the only exchange is one
16-byte {best f, edge index} record per rank (the batched form of the incumbent prune of
SamplingBasedPlanner::pushVertexQueue, SamplingBasedPlanner.cpp:11-13), all-gathered with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).  No arithmetic of the path lives here."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous [lo, hi) of `n` edges for `rank`; sizes differ by at most one.  Callers keep edges
    grouped by source vertex so a parent's ribbon set is staged by neighbouring warps."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scenario_assignment(n_scenarios, rank, world_size):
    """Independent planning scenarios are dealt round-robin: scenario s -> rank s mod world_size."""
    return list(range(rank, int(n_scenarios), int(world_size)))


def pack_best(f, local_index, lo):
    """{f64 f; i64 global edge index} as two float64 lanes of raw bits (what ppe_best_copy_device writes,
    shifted to the global index).  local_index < 0 means "no feasible edge"."""
    rec = np.zeros(2, dtype=np.float64)
    rec[0] = f if local_index >= 0 else np.inf
    rec.view(np.int64)[1] = (lo + local_index) if local_index >= 0 else -1
    return rec


def gather_best(local_record, group=None):
    """All-gather of one 16-byte record per rank.  `local_record`: float64[2] tensor (CPU for gloo, CUDA
    for NCCL) whose second lane holds the int64 bits of the global edge index.  Returns
    (best f, global edge index, owning rank); ties go to the smaller edge index, as K3 does on one GPU."""
    world_size = dist.get_world_size(group)
    out = torch.empty(2 * world_size, dtype=torch.float64, device=local_record.device)
    dist.all_gather_into_tensor(out, local_record.contiguous(), group=group)
    rec = out.cpu().numpy().reshape(world_size, 2)
    fs = rec[:, 0].copy()
    idx = rec[:, 1].copy().view(np.int64)
    best = (np.inf, -1, -1)
    for r in range(world_size):
        if idx[r] < 0:
            continue
        if best[1] < 0 or fs[r] < best[0] or (fs[r] == best[0] and idx[r] < best[1]):
            best = (float(fs[r]), int(idx[r]), r)
    return best
