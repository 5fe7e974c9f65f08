"""Ring / look-ahead / tiles-per-CTA sweep of the production conv kernel on the layer shapes that dominate the bench step,
with the neighbour-pattern row order active (us3d_debug_set_tuning overrides the launcher's choices)."""
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
raw.us3d_debug_set_tuning.argtypes = [ctypes.c_int] * 3
raw.us3d_debug_set_tuning.restype = None
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key1 = x0.coordinate_manager, x0.coordinate_map_key
key2 = cm.stride(key1, (2, 2, 2))
key4 = cm.stride(key2, (2, 2, 2))
tables = {1: cm.forward_table(key1, key1, (3, 3, 3)), 2: cm.forward_table(key2, key2, (3, 3, 3)), 4: cm.forward_table(key4, key4, (3, 3, 3))}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(level, cin, cout, a_slots, lag, T):
    table = tables[level]
    x = torch.randn(table.n_rows, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    raw.us3d_debug_set_tuning(a_slots, lag, T)
    for _ in range(2):
        Fn.spconv_gather(x, table, w, cin, cout, False, False)
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Fn.spconv_gather(x, table, w, cin, cout, False, False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[2]


for level, cin, cout in ((1, 96, 96), (1, 128, 96), (2, 96, 96), (2, 32, 32), (4, 64, 64), (4, 128, 128)):
    base = run(level, cin, cout, 0, 0, 0)
    line = [f"L{level} {cin}->{cout}: auto {base:.3f}"]
    for lag in (1, 2, 3):
        for T in (1, 2, 4):
            line.append(f"lag{lag}/T{T} {run(level, cin, cout, 0, lag, T):.3f}")
    print(" | ".join(line), flush=True)
raw.us3d_debug_set_tuning(0, 0, 0)
