"""GPU timeline + host profile of the C3 step (Mask3D self-training step, 4 x 200k-voxel scenes)."""
import os, sys, time, collections, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench_configs as B
from unscene3d_b200 import engine

dev = torch.device("cuda")
engine.set_coordinate_stream(torch.cuda.Stream(device=dev, priority=-1), dev)
net, crit, wd = B.build_mask3d(dev)
batch = B.scene_batch(int(os.environ.get("SCENES", "4")), 200_000, 100, dev)
step = lambda: B.train_step(net, crit, wd, batch, 1)
for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host {1e3 * (t1 - t0):.1f} ms, with drain {1e3 * (t2 - t0):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
busy = 0.0; cur_s = cur_e = None
for s, e, name in ks:
    if cur_e is None: cur_s, cur_e = s, e
    elif s <= cur_e: cur_e = max(cur_e, e)
    else: busy += cur_e - cur_s; cur_s, cur_e = s, e
busy += cur_e - cur_s
print(f"GPU span {(ks[-1][1] - ks[0][0]) / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, {len(ks)} kernels+memops")
tot = collections.defaultdict(lambda: [0, 0.0])
for s, e, name in ks:
    tot[name[:90]][0] += 1; tot[name[:90]][1] += e - s
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {t / 1e3:8.3f} ms {n:6d}  {name}")
pr = cProfile.Profile(); pr.enable(); step(); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
