"""world_size-2 gloo tests (CPU) of the data-parallel host logic: scene sharding, max-over-ranks timing reduction and
the bucketed gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unscene3d_b200 import distributed as D

    assert D.world_info() == (rank, world)
    seeds = [D.scene_seed(100, step, rank, world, scenes_per_rank=2, slot=s) for step in range(3) for s in range(2)]
    ms = D.max_over_ranks(10.0 + rank)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(64, 256), torch.nn.ReLU(), torch.nn.Linear(256, 8))
    x = torch.full((4, 64), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    D.allreduce_gradients(net.parameters(), bucket_bytes=1 << 12)  # several buckets
    out[rank] = (seeds, ms, local, [p.grad.clone() for p in net.parameters()])
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (s0, ms0, l0, g0), (s1, ms1, l1, g1) = out[0], out[1]
    assert not set(s0) & set(s1) and len(set(s0 + s1)) == 12, "ranks must see disjoint scenes"
    assert ms0 == ms1 == 11.0, "timings are reduced with MAX over ranks"
    for a, b, la, lb in zip(g0, g1, l0, l1):
        assert torch.allclose(a, b) and torch.allclose(a, (la + lb) / 2, atol=1e-6)
