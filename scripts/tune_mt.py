"""Where does the multi-tile conv kernel wait?  Times one big layer under ring/lag/T overrides and prints the MMA
thread's wait-cycle breakdown (us3d_debug_set_prof / us3d_debug_set_tuning)."""
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
raw.us3d_debug_set_tuning.argtypes = [ctypes.c_int] * 3
raw.us3d_debug_set_tuning.restype = None

dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
table = cm.forward_table(key, key, (3, 3, 3))
prof = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
Fn._tc_kernel["fwd"] = "mt"


def run(cin, cout, mode, a_slots, lag, T, label=""):
    x = torch.randn(s.n, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    Fn.set_precision(mode)
    raw.us3d_debug_set_tuning(a_slots, lag, T)
    raw.us3d_debug_set_prof(ctypes.c_void_p(prof.data_ptr()))
    for _ in range(2):
        Fn.spconv_gather(x, table, w, cin, cout, False, False)
    torch.cuda.synchronize()
    prof.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    Fn.spconv_gather(x, table, w, cin, cout, False, False)
    b.record()
    torch.cuda.synchronize()
    p = prof.view(148, 8).double().cpu()
    busy = p[p[:, 0] > 0]
    tot, wacc, wb, wa, items, bitems = [float(busy[:, i].mean()) for i in range(6)]
    print(f"{cin}->{cout} mode {mode} a_slots={a_slots} lag={lag} T={T} {label}: {a.elapsed_time(b):.3f} ms | MMA thread: total {tot:.0f} cyc, "
          f"wait acc {wacc / tot * 100:.0f}% b {wb / tot * 100:.0f}% a {wa / tot * 100:.0f}% | per A-item {tot / max(items, 1):.0f} cyc "
          f"(items {items:.0f}, b-items {bitems:.0f})", flush=True)


for mode in (1, 3):
    run(128, 96, mode, 0, 0, 0, "auto")
    run(128, 96, mode, 0, 1, 0)
    run(128, 96, mode, 0, 2, 0)
    run(128, 96, mode, 2, 1, 0)
    run(128, 96, mode, 0, 0, 1)
    run(128, 96, mode, 0, 0, 2)
run(64, 64, 3, 0, 0, 0)
run(256, 256, 3, 0, 0, 0)
raw.us3d_debug_set_tuning(0, 0, 0)
raw.us3d_debug_set_prof(None)
