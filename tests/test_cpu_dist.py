"""CPU: the N>1 host logic (sharding + the single all-gather) over gloo with world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omg_planner_b200 import dist as D


def test_shard_range_covers_batch():
    for batch in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, batch):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = torch.arange(batch, dtype=torch.float64) * 1.5
    lo, hi = D.shard_range(batch, rank, world)
    got = D.all_gather_ragged(costs[lo:hi].clone(), batch)
    assert torch.equal(got, costs)
    eq = D.all_gather_costs(torch.full((4, 2), float(rank), dtype=torch.float64))
    assert eq.shape == (8, 2) and torch.equal(eq[:4], torch.zeros(4, 2, dtype=torch.float64)) and (eq[4:] == 1).all()
    dist.destroy_process_group()


def test_all_gather_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, 9), nprocs=2, join=True)


def test_numa_binding_helpers():
    assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert D._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert D.bind_host_to_device_numa(0) is None or isinstance(D.bind_host_to_device_numa(0), dict)   # no GPU here: no-op
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
