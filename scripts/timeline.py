"""GPU timeline of bench steps through torch.profiler (CUPTI): device busy time vs step wall time, per-kernel totals and
the idle gaps of the stream (who starves the GPU).  nsys is not installed in this image; this is the substitute."""
import os, sys, time, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200
from unscene3d_b200 import engine, models
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene
from unscene3d_b200.utils import BackboneConfig, seeded_state
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
scene = make_scene(int(os.environ.get("US3D_VOXELS", "200000")), seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((scene.n, 1), np.int32), scene.coords], 1)).to(dev)
f = torch.from_numpy(scene.colors).to(dev)
net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
net.load_state_dict(seeded_state(net, 0))
net = net.to(dev).train()
w = torch.linspace(-1, 1, 96, device=dev)

def step():
    Fn.invalidate_packed_weights()
    x = engine.SparseTensor(f, c4)
    out, _ = net(x)
    loss = (out.F * w).mean()
    loss.backward()
    net.zero_grad(set_to_none=True)

for _ in range(4):
    step()
torch.cuda.synchronize()
STEPS = 3
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / STEPS * 1e3
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
busy = 0.0; cur_s, cur_e = None, None
gaps = []
for s, e, name in ks:
    if cur_e is None:
        cur_s, cur_e = s, e
    elif s <= cur_e:
        cur_e = max(cur_e, e)
    else:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, prev_name, name))
        cur_s, cur_e = s, e
    prev_name = name
busy += cur_e - cur_s
span = ks[-1][1] - ks[0][0]
print(f"wall/step {wall:.2f} ms (under profiler) | GPU span {span / STEPS / 1e3:.2f} ms/step, GPU busy {busy / STEPS / 1e3:.2f} ms/step, idle {(span - busy) / STEPS / 1e3:.2f} ms/step, {len(ks) / STEPS:.0f} kernels+memops/step")
tot = collections.defaultdict(lambda: [0, 0.0])
for s, e, name in ks:
    tot[name[:70]][0] += 1; tot[name[:70]][1] += e - s
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"  {t / STEPS / 1e3:8.3f} ms/step {n / STEPS:7.1f}/step  {name}")
hist = collections.Counter()
for g, a, b in gaps:
    hist[min(int(g // 10) * 10, 200)] += 1
print("idle gap histogram (us -> count/step):", {k: round(v / STEPS, 1) for k, v in sorted(hist.items())})
big = sorted(gaps, key=lambda t: -t[0])[:12]
for g, a, b in big:
    print(f"  gap {g:8.1f} us after {a[:50]} before {b[:50]}")
