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


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_device_numa(device_index):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    host-buffer entry points (first touch) and the staging memcpys are node-local.  One process per GPU drives
    ~5 MB per step over PCIe in each direction of the link; with 8 processes on a two-socket box, buffers that land
    on the far socket cross the inter-socket link as well.  A no-op (returns None) when the topology cannot be read,
    the node has no CPUs of the current affinity mask, or there is a single node.  Returns a small dict otherwise."""
    import os

    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0 or not os.path.isdir("/sys/devices/system/node/node1"):
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read()) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "numa_node": node, "cpus": len(cpus)}
    except Exception:   # noqa: BLE001 -- best effort: any failure leaves the affinity as it was
        return None
