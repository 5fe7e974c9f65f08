"""Round-2 sweep of the production conv kernel (us3d_spconv_gather_mt): producer completion (wait_group look-ahead 1 / 2 vs
cp.async.mbarrier.arrive.noinc), fused [W_hi | W_lo] operand on / off, on every level of the 200k-voxel bench scene, L2 flushed
and L2 warm, each variant checked against the exact-fp32 SIMT kernel."""
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
OLD = not hasattr(raw, "us3d_debug_set_tuning4")  # round-1 library (A/B through US3D_LIB)
if OLD:
    raw.us3d_debug_set_tuning.argtypes = [ctypes.c_int] * 3
    raw.us3d_debug_set_tuning.restype = None
    raw.us3d_debug_set_tuning4 = lambda a, lag, T, fuse: raw.us3d_debug_set_tuning(a, lag, T)
else:
    raw.us3d_debug_set_tuning4.argtypes = [ctypes.c_int] * 4
    raw.us3d_debug_set_tuning4.restype = None
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"],
                     capture_output=True, text=True).stdout.strip(), flush=True)
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
keys = [key]
for _ in range(4):
    keys.append(cm.stride(keys[-1], (2, 2, 2)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(table, x, w, cin, cout, lag, fuse, T=0, do_flush=True, reps=5):
    raw.us3d_debug_set_tuning4(0, lag, T, fuse)
    for _ in range(2):
        y = Fn.spconv_gather(x, table, w, cin, cout, False, False)
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Fn.spconv_gather(x, table, w, cin, cout, False, False)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    raw.us3d_debug_set_tuning4(0, 0, 0, 0)
    return sorted(ts)[len(ts) // 2], y


shapes = [(0, 96, 96), (0, 128, 96), (1, 96, 96), (1, 32, 32), (2, 64, 64), (2, 128, 128), (3, 128, 128), (3, 256, 256), (4, 256, 256)]
only = os.environ.get("US3D_LEVELS")
for lvl, cin, cout in shapes:
    if only and str(lvl) not in only.split(","):
        continue
    k = keys[lvl]
    table = cm.forward_table(k, k, (3, 3, 3))
    n = table.n_rows
    g = torch.Generator(device=dev).manual_seed(lvl * 100 + cin)
    x = torch.randn(n, cin, device=dev, generator=g)
    w = torch.randn(27, cin, cout, device=dev, generator=g) * 0.03
    Fn.set_precision(0)
    ref = Fn.spconv_gather(x, table, w, cin, cout, False, False).double()
    Fn.set_precision(3)
    line = [f"L{lvl} n={n} {cin}->{cout}:"]
    for lag in ((1, 2) if OLD else (1, 2, 9)):
        for fuse in ((1,) if OLD else (1, 2)):
            if fuse == 2 and 2 * cout > 256:
                continue
            t_cold, y = run(table, x, w, cin, cout, lag, fuse)
            t_warm, _ = run(table, x, w, cin, cout, lag, fuse, do_flush=False)
            err = float((y.double() - ref).norm() / ref.norm())
            line.append(f"lag{lag}/{'fuse' if fuse == 2 else 'sep '} {t_cold * 1e3:6.1f}/{t_warm * 1e3:6.1f}us e={err:.1e}")
    print(" | ".join(line), flush=True)
if OLD:
    sys.exit(0)
# tiles per CTA at the two fine levels with the fused operand (accumulators are twice as wide: T <= 2 at Cout 96)
for lvl, cin, cout in ((0, 96, 96), (1, 96, 96), (1, 32, 32)):
    k = keys[lvl]
    table = cm.forward_table(k, k, (3, 3, 3))
    x = torch.randn(table.n_rows, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    line = [f"L{lvl} {cin}->{cout} T sweep (lag 9):"]
    for fuse in (1, 2):
        for T in (1, 2, 4):
            t, _ = run(table, x, w, cin, cout, 9, fuse, T)
            line.append(f"{'fuse' if fuse == 2 else 'sep '} T{T} {t * 1e3:6.1f}us")
    print(" | ".join(line), flush=True)
# natural row order for comparison at the finest level
engine.set_row_ordering(0)
k = keys[0]
t0 = cm.forward_table(k, k, (3, 3, 3))
table = engine.NeighbourTable(t0.nbr, t0.mask, t0.n_rows, t0.kvol)
x = torch.randn(table.n_rows, 96, device=dev)
w = torch.randn(27, 96, 96, device=dev) * 0.03
for lag in (1, 9):
    for fuse in (1, 2):
        t, _ = run(table, x, w, 96, 96, lag, fuse)
        print(f"natural order L0 96->96 lag{lag} {'fuse' if fuse == 2 else 'sep '}: {t * 1e3:.1f} us", flush=True)
