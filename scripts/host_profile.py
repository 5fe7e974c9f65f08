"""cProfile of the host side of one bench step (where does the Python time go?)."""
import cProfile, os, pstats, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200
from unscene3d_b200 import engine, models
from unscene3d_b200.synthetic import make_scene
from unscene3d_b200.utils import BackboneConfig, seeded_state

dev = torch.device("cuda")
scene = make_scene(int(os.environ.get("US3D_VOXELS", "200000")), seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((scene.n, 1), np.int32), scene.coords], 1)).to(dev)
f = torch.from_numpy(scene.colors).to(dev)
net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
net.load_state_dict(seeded_state(net, 0))
net = net.to(dev).train()
w = torch.linspace(-1, 1, 96, device=dev)

def step():
    x = engine.SparseTensor(f, c4)
    out, _ = net(x)
    loss = (out.F * w).mean()
    loss.backward()
    net.zero_grad(set_to_none=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time/step {(t1 - t) / 5 * 1e3:.2f} ms, incl. drain {(t2 - t) / 5 * 1e3:.2f} ms")
class _TimedLib:
    """Proxy over the ctypes library that accumulates host wall time per entry point."""
    def __init__(self, lib):
        self._lib, self.acc, self.samples = lib, {}, {}
    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        acc = self.acc.setdefault(name, [0, 0.0])
        samples = self.samples.setdefault(name, [])
        def timed(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            dt = time.perf_counter() - t0
            acc[0] += 1; acc[1] += dt
            samples.append((dt, threading.get_ident(), a))
            return r
        return timed

from unscene3d_b200.engine import functional as Fn, coords as Co
tl = _TimedLib(Fn.lib)
Fn.lib = tl; Co.lib = tl
t = time.perf_counter()
for _ in range(3):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host time/step with timed lib {(t1 - t) / 3 * 1e3:.2f} ms; C-ABI calls:")
for name, (n, tot) in sorted(tl.acc.items(), key=lambda kv: -kv[1][1]):
    print(f"  {name:34s} {n / 3:7.1f} calls/step  {tot / 3 * 1e3:8.3f} ms/step  {tot / max(n, 1) * 1e6:8.1f} us/call")
sm = tl.samples.get("us3d_spconv_gather_mt_bn", [])
main = threading.get_ident()
for label, sel in (("main thread", [x for x in sm if x[1] == main]), ("other threads", [x for x in sm if x[1] != main])):
    d = sorted(x[0] * 1e6 for x in sel)
    if d:
        print(f"  gather_mt {label}: n={len(d)} min {d[0]:.1f} median {d[len(d) // 2]:.1f} p90 {d[int(len(d) * 0.9)]:.1f} max {d[-1]:.1f} us")
for x in sorted(sm, key=lambda x: -x[0])[:12]:
    a = x[2]
    print(f"    {x[0] * 1e6:8.1f} us  thread {'main' if x[1] == main else 'bwd'}  n_in {a[2]} n_rows {a[4]} kvol {a[5]} cin {a[7]} cout {a[8]} acc {a[14]}")
for x in sorted(sm, key=lambda x: x[0])[:6]:
    a = x[2]
    print(f"    {x[0] * 1e6:8.1f} us  thread {'main' if x[1] == main else 'bwd'}  n_in {a[2]} n_rows {a[4]} kvol {a[5]} cin {a[7]} cout {a[8]} acc {a[14]}")
Fn.lib = tl._lib; Co.lib = tl._lib
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(40)
