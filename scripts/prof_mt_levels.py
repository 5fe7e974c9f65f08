"""MMA-thread view of the production conv kernel on every level of the 200k-voxel bench scene: kernel time, cycles of the
busiest CTA, share of the MMA thread's time spent waiting for the accumulator / weight slabs / gathered tiles, cycles until
the first gathered tile landed (us3d_debug_set_prof)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import unscene3d_b200  # noqa: F401
from unscene3d_b200 import _lib, engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene

raw = ctypes.CDLL(_lib.LIB_PATH)
raw.us3d_debug_set_prof.argtypes = [ctypes.c_void_p]
raw.us3d_debug_set_prof.restype = None
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
keys = [key]
for _ in range(4):
    keys.append(cm.stride(keys[-1], (2, 2, 2)))
prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

shapes = [(0, 96, 96), (1, 96, 96), (1, 32, 32), (2, 64, 64), (2, 128, 128), (3, 128, 128), (3, 256, 256), (4, 256, 256)]
for lvl, cin, cout in shapes:
    k = keys[lvl]
    table = cm.forward_table(k, k, (3, 3, 3))
    x = torch.randn(table.n_rows, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    for _ in range(2):
        Fn.spconv_gather(x, table, w, cin, cout, False, False)
    torch.cuda.synchronize()
    raw.us3d_debug_set_prof(ctypes.c_void_p(prof.data_ptr()))
    prof.zero_()
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    Fn.spconv_gather(x, table, w, cin, cout, False, False)
    b.record()
    torch.cuda.synchronize()
    raw.us3d_debug_set_prof(None)
    p = prof.view(148, 8).double().cpu()
    busy = p[p[:, 0] > 0]
    tot = busy[:, 0]
    i = int(tot.argmax())
    r = busy[i]
    print(f"L{lvl} n={table.n_rows} {cin}->{cout}: {a.elapsed_time(b) * 1e3:.1f} us | CTAs {busy.shape[0]} | busiest CTA {r[0]:.0f} cyc "
          f"({r[0] / 1.965e3:.1f} us @1.965 GHz): wait acc {r[1] / r[0] * 100:.0f}% b {r[2] / r[0] * 100:.0f}% a {r[3] / r[0] * 100:.0f}% | items {r[4]:.0f} "
          f"b-items {r[5]:.0f} | first tile after {r[6]:.0f} cyc | mean CTA {float(tot.mean()):.0f} cyc", flush=True)
