"""Per-layer timing of the production sparse-conv kernels (forward/dgrad gather kernel and weight gradient) on the
coordinate maps of the 200k-voxel bench scene.  CUDA events, median of 7 launches each, L2-warm (relative tuning only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene

dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, k1 = x0.coordinate_manager, x0.coordinate_map_key
keys = [k1]
for _ in range(4):
    keys.append(cm.stride(keys[-1], (2, 2, 2)))
mode = int(os.environ.get("US3D_MODE", "3"))
Fn.set_precision(mode)


def timeit(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts))


shapes = [(0, 96, 96), (0, 128, 96), (0, 32, 32), (1, 96, 96), (1, 32, 32), (1, 128, 96), (2, 64, 64), (2, 128, 128), (3, 128, 128),
          (3, 256, 256), (4, 256, 256)]
only = os.environ.get("US3D_SHAPES")
for lvl, cin, cout in shapes:
    key = keys[lvl]
    n = cm.size(key)
    table = cm.forward_table(key, key, (3, 3, 3))
    x = torch.randn(n, cin, device=dev)
    dy = torch.randn(n, cout, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    wp = Fn.pack_weights(w, False, False, mode)
    hi, lo = Fn.bf16_planes(x, mode == 3)
    dh, dl = Fn.bf16_planes(dy, mode == 3)
    t_f = timeit(lambda: Fn.spconv_gather(x, table, w, cin, cout, False, False))
    t_w = timeit(lambda: Fn.spconv_wgrad(x, table, dy, cin, cout))
    fl = 2.0 * n * 27 * cin * cout
    print(f"s{2**lvl:<2d} n={n:6d} {cin:3d}->{cout:3d}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} dense-equiv TFLOP/s)   wgrad {t_w:7.1f} us", flush=True)
