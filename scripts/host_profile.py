"""cProfile of the host side of one bench step (where does the Python time go?)."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200
from unscene3d_b200 import engine, models
from unscene3d_b200.synthetic import make_scene
from unscene3d_b200.utils import BackboneConfig, seeded_state

dev = torch.device("cuda")
scene = make_scene(200_000, seed=0, with_masks=False)
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
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
