"""Multi-GPU plumbing for the one place the path shards: trajectories are independent, so the
batch is split contiguously across ranks, the scene is replicated, and the only exchange is ONE all-gather of
the final per-trajectory costs (optionally termination flags / trajectories) at the end of a plan."""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous [lo, hi) of a batch of `batch` trajectories owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_costs(local, group=None):
    """All-gather equal-sized per-trajectory tensors [B_local, ...] -> [world*B_local, ...] on every rank.
    NCCL on GPUs (NVLink/NVSwitch; a few KB, latency-bound), gloo in the CPU tests."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def all_gather_ragged(local, batch, group=None):
    """All-gather shards produced by shard_range (sizes may differ by one) back into batch order."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: sizes[rank]] = local
    out = all_gather_costs(pad, group).reshape((world, mx) + tuple(local.shape[1:]))
    return torch.cat([out[r, : sizes[r]] for r in range(world)], 0)
