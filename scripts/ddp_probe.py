"""Two-rank probe of GradientReducer over NCCL on a small torch model (no libus3d): hooks fire all-reduces from the autograd
thread while backward runs; checks the averaged gradients against a manual all-reduce."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import faulthandler

import torch
import torch.distributed as dist

faulthandler.dump_traceback_later(60, exit=True)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from unscene3d_b200 import distributed as D

torch.manual_seed(0)
net = torch.nn.Sequential(*[torch.nn.Linear(2048, 2048) for _ in range(8)]).to(dev)
red = D.GradientReducer(net.parameters(), bucket_bytes=32 << 20)
print(f"rank {rank}: {len(red.buckets)} buckets, {red.total_bytes / 1e6:.0f} MB", flush=True)
for step in range(5):
    x = torch.full((64, 2048), float(rank + 1 + step), device=dev)
    net(x).sum().backward()
    red.finish()
    g = [p.grad.clone() for p in net.parameters()]
    red.zero_grad()
    torch.cuda.synchronize()
    print(f"rank {rank}: step {step} ok, |g0| {float(g[0].norm()):.4f}", flush=True)
dist.barrier()
dist.destroy_process_group()
print(f"rank {rank}: done", flush=True)
