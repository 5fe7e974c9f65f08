"""Inclusive host wall time of the step's main Python functions on a tiny scene (no GPU back-pressure): where do the ~14 ms of
host time per step go?  (perf_counter wrappers; cProfile's per-call overhead distorts a step of 26k short calls.)"""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200
from unscene3d_b200 import engine, models
from unscene3d_b200.engine import functional as Fn, blocks as B, coords as Co, tensor as T
from unscene3d_b200.synthetic import make_scene
from unscene3d_b200.utils import BackboneConfig, seeded_state

dev = torch.device("cuda")
scene = make_scene(int(os.environ.get("US3D_VOXELS", "3000")), seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((scene.n, 1), np.int32), scene.coords], 1)).to(dev)
f = torch.from_numpy(scene.colors).to(dev)
net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
net.load_state_dict(seeded_state(net, 0))
net = net.to(dev).train()
w = torch.linspace(-1, 1, 96, device=dev)
acc = collections.defaultdict(lambda: [0, 0.0])

def wrap(obj, name, label=None):
    fn = getattr(obj, name)
    label = label or name
    def timed(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            e = acc[label]; e[0] += 1; e[1] += time.perf_counter() - t0
    setattr(obj, name, timed)

def step():
    Fn.pack_network(net)
    x = engine.SparseTensor(f, c4)
    out, _ = net(x)
    loss = (out.F * w).mean()
    loss.backward()
    net.zero_grad(set_to_none=True)

for _ in range(5):
    step()
torch.cuda.synchronize()
N = 10
t0 = time.perf_counter()
for _ in range(N):
    step()
base = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
print(f"host ms/step (no instrumentation): {base:.2f}")
for name in ("spconv_gather", "spconv_wgrad", "conv_input_gradient", "bn_apply_raw", "bn_backward_raw", "bn_batch_stats", "pack_network", "packed_weights", "bf16_planes"):
    wrap(Fn, name)
wrap(B.FusedBasicBlockFunction, "forward", "block.forward"); wrap(B.FusedBasicBlockFunction, "backward", "block.backward")
wrap(B.FusedConvNormReLUFunction, "forward", "cnr.forward"); wrap(B.FusedConvNormReLUFunction, "backward", "cnr.backward")
wrap(B, "fused_basic_block"); wrap(B, "fused_conv_norm_relu")
wrap(torch, "empty"); wrap(torch, "zeros"); wrap(torch, "empty_like")
wrap(engine, "SparseTensor", "SparseTensor()")
samples = {}
class TL:
    def __init__(self, lib): self._lib = lib; self._c = {}
    def __getattr__(self, name):
        fn = self._c.get(name)
        if fn is None:
            real = getattr(self._lib, name)
            def fn(*a, _real=real, _e=acc["C:" + name], _pc=time.perf_counter, _s=samples.setdefault(name, [])):
                t0 = _pc(); r = _real(*a); dt = _pc() - t0; _e[0] += 1; _e[1] += dt; _s.append((dt, a)); return r
            self._c[name] = fn
        return fn
tl = TL(Fn.lib); Fn.lib = tl; Co.lib = tl; B.lib = tl
t0 = time.perf_counter()
for _ in range(N):
    step()
inst = (time.perf_counter() - t0) / N * 1e3
torch.cuda.synchronize()
print(f"host ms/step (instrumented): {inst:.2f}")
csum = sum(v[1] for k, v in acc.items() if k.startswith("C:")) / N * 1e3
print(f"  all C-ABI calls: {csum:.2f} ms/step, {sum(v[0] for k, v in acc.items() if k.startswith('C:')) / N:.0f} calls/step")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"  {k:34s} {v[0] / N:7.1f} calls/step {v[1] / N * 1e3:8.3f} ms/step {v[1] / max(v[0], 1) * 1e6:8.1f} us/call")
sm = samples.get("us3d_spconv_gather_mt_bn", [])
d = sorted(x[0] * 1e6 for x in sm)
print(f"gather_mt_bn: n={len(d)} min {d[0]:.1f} p10 {d[len(d) // 10]:.1f} median {d[len(d) // 2]:.1f} p90 {d[int(len(d) * 0.9)]:.1f} p99 {d[int(len(d) * 0.99)]:.1f} max {d[-1]:.1f} us")
for x in sorted(sm, key=lambda x: -x[0])[:8]:
    a = x[1]
    print(f"    {x[0] * 1e6:8.1f} us  n_in {a[2]} n_rows {a[4]} kvol {a[5]} cin {a[7]} cout {a[8]} bn {a[19] is not None}")
by = collections.defaultdict(list)
for dt, a in sm:
    by[(a[4], a[5], a[7], a[8], a[19] is not None)].append(dt * 1e6)
for k, v in sorted(by.items(), key=lambda kv: -sum(kv[1]))[:14]:
    print(f"    n_rows {k[0]:6d} kvol {k[1]:2d} {k[2]:3d}->{k[3]:3d} bn {k[4]!s:5s}: n={len(v):3d} mean {sum(v) / len(v):7.1f} us  min {min(v):6.1f}")
# forward / backward split
def fwd_only():
    x = engine.SparseTensor(f, c4)
    out, _ = net(x)
    return (out.F * w).mean()
ts = []
for _ in range(N):
    Fn.pack_network(net)
    t0 = time.perf_counter(); loss = fwd_only(); t1 = time.perf_counter(); loss.backward(); t2 = time.perf_counter()
    net.zero_grad(set_to_none=True); t3 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2))
ts = np.array(ts) * 1e3
print("forward / backward / zero_grad ms (instrumented):", ts.mean(0).round(2))
