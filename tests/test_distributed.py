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
    reduced = [p.grad.clone() for p in net.parameters()]
    # ---- GradientReducer: overlapped buckets, both ways of filling them; rank 1 does not use the `extra` branch, and the
    # model mixes dtypes (buckets are per dtype, the parameter list — not "who has a gradient" — defines them)
    red_out = {}
    for as_views in (False, True):
        torch.manual_seed(1)
        m = torch.nn.ModuleDict({"a": torch.nn.Linear(32, 64), "extra": torch.nn.Linear(64, 64), "b": torch.nn.Linear(64, 4),
                                 "h": torch.nn.Linear(4, 4).to(torch.float64)})
        red = D.GradientReducer(m.parameters(), bucket_bytes=1 << 11, as_views=as_views)
        assert len(red.buckets) >= 3 and len({b.flat.dtype for b in red.buckets}) == 2
        res = []
        for step in range(2):
            xx = torch.full((3, 32), float(rank + 1 + step))
            hid = torch.relu(m["a"](xx))
            if rank == 0:
                hid = hid + m["extra"](hid)
            y = m["h"](m["b"](hid).double())
            y.sum().backward()
            red.finish()
            assert all(p.grad is not None for p in m.parameters())
            assert all(p.grad.data_ptr() == v.data_ptr() for b in red.buckets for p, v in zip(b.params, b.views))
            res.append({k: p.grad.clone() for k, p in m.named_parameters()})
            red.zero_grad()
        red.remove()
        red_out[as_views] = res
    out[rank] = (seeds, ms, local, reduced, red_out)
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    (s0, ms0, l0, g0, r0), (s1, ms1, l1, g1, r1) = out[0], out[1]
    assert not set(s0) & set(s1) and len(set(s0 + s1)) == 12, "ranks must see disjoint scenes"
    assert ms0 == ms1 == 11.0, "timings are reduced with MAX over ranks"
    for a, b, la, lb in zip(g0, g1, l0, l1):
        assert torch.allclose(a, b) and torch.allclose(a, (la + lb) / 2, atol=1e-6)
    # GradientReducer: identical averaged gradients on both ranks, identical between the two bucket-filling modes, and equal
    # to the average of single-process gradients (the unused branch contributes zeros on rank 1)
    for as_views in (False, True):
        for step in range(2):
            a, b = r0[as_views][step], r1[as_views][step]
            assert a.keys() == b.keys()
            for k in a:
                assert torch.equal(a[k], b[k]), (as_views, step, k)
                assert torch.allclose(a[k], r0[False][step][k], atol=1e-7)
    want = _single_process_reference()
    for step in range(2):
        for k, v in want[step].items():
            assert torch.allclose(r0[False][step][k], v, atol=1e-6, rtol=1e-5), (step, k)


def _single_process_reference():
    res = []
    for step in range(2):
        grads = []
        for rank in range(2):
            torch.manual_seed(1)
            m = torch.nn.ModuleDict({"a": torch.nn.Linear(32, 64), "extra": torch.nn.Linear(64, 64), "b": torch.nn.Linear(64, 4),
                                     "h": torch.nn.Linear(4, 4).to(torch.float64)})
            xx = torch.full((3, 32), float(rank + 1 + step))
            hid = torch.relu(m["a"](xx))
            if rank == 0:
                hid = hid + m["extra"](hid)
            m["h"](m["b"](hid).double()).sum().backward()
            grads.append({k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()})
        res.append({k: (grads[0][k] + grads[1][k]) / 2 for k in grads[0]})
    return res


def test_gradient_reducer_single_process_keeps_gradients_in_buckets():
    from unscene3d_b200 import distributed as D

    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 2))
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 2))
    ref.load_state_dict(net.state_dict())
    x = torch.randn(5, 16)
    ref(x).sum().backward()
    for as_views in (False, True):
        red = D.GradientReducer(net.parameters(), bucket_bytes=256, as_views=as_views)
        net(x).sum().backward()
        red.finish()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.equal(p.grad, q.grad)
        assert red.total_bytes == sum(p.numel() * 4 for p in net.parameters())
        red.zero_grad()
        red.remove()
        net.zero_grad(set_to_none=True)
