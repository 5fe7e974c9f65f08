"""Role view of the weight-gradient kernel on the levels of the 200k-voxel bench scene (us3d_debug_set_prof_wgrad): kernel time,
the busiest CTA's producer loop / MMA loop in cycles and the share of each spent waiting."""
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
raw.us3d_debug_set_prof_wgrad.argtypes = [ctypes.c_void_p]
raw.us3d_debug_set_prof_wgrad.restype = None
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
keys = [key]
for _ in range(4):
    keys.append(cm.stride(keys[-1], (2, 2, 2)))
prof = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(0, 96, 96), (0, 32, 32), (1, 96, 96), (1, 32, 32), (2, 64, 64), (2, 128, 128), (3, 256, 256)]
for lvl, cin, cout in shapes:
    k = keys[lvl]
    table = cm.forward_table(k, k, (3, 3, 3))
    n = table.n_rows
    x = torch.randn(n, cin, device=dev)
    dy = torch.randn(n, cout, device=dev)
    for _ in range(2):
        Fn.spconv_wgrad(x, table, dy, cin, cout)
    torch.cuda.synchronize()
    raw.us3d_debug_set_prof_wgrad(ctypes.c_void_p(prof.data_ptr()))
    prof.zero_()
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    Fn.spconv_wgrad(x, table, dy, cin, cout)
    b.record()
    torch.cuda.synchronize()
    raw.us3d_debug_set_prof_wgrad(None)
    p = prof.view(-1, 16).double().cpu()
    p = p[p[:, 4] > 0]
    i = int(p[:, 4].argmax())
    r = p[i]
    print(f"L{lvl} n={n} {cin}->{cout}: {a.elapsed_time(b) * 1e3:.1f} us incl. plane split | CTAs {p.shape[0]} | producers {r[0]:.0f} cyc "
          f"(wait dY slot {r[1] / r[0] * 100:.0f}% X slot {r[2] / r[0] * 100:.0f}%) epilogue {r[3]:.0f} | MMA warp {r[4]:.0f} cyc "
          f"(wait dY {r[5] / r[4] * 100:.0f}% X {r[6] / r[4] * 100:.0f}% issue {r[7] / r[4] * 100:.0f}%) slots {r[8]:.0f} -> {r[4] / max(r[8], 1):.0f} cyc/slot"
          f" | mean MMA loop {float(p[:, 4].mean()):.0f}", flush=True)
