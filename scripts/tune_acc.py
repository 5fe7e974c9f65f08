"""A/B of the conv kernel's accumulator layout on the levels of the 200k-voxel scene: fused [W_hi | W_lo] operand with one
accumulator set (T = 2), unfused with two sets (T = 2), fused with two sets (T = 1).  In-library timing is not used: CUDA events
around the call, planes and weight image prepared outside, L2 flushed."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import _lib, engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene

raw = ctypes.CDLL(_lib.LIB_PATH)
raw.us3d_debug_set_tuning4.argtypes = [ctypes.c_int] * 4
raw.us3d_debug_set_tuning_acc.argtypes = [ctypes.c_int]
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
keys = [key]
for _ in range(4):
    keys.append(cm.stride(keys[-1], (2, 2, 2)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(0, 96, 96), (0, 128, 96), (1, 96, 96), (1, 32, 32), (2, 64, 64), (2, 128, 128), (3, 128, 128), (3, 256, 256), (4, 256, 256)]
variants = [("fuse,1set", 2, 1), ("nofuse,2set", 1, 2), ("fuse,2set", 2, 2), ("nofuse,1set", 1, 1)]
for lvl, cin, cout in shapes:
    table = cm.forward_table(keys[lvl], keys[lvl], (3, 3, 3))
    n = table.n_rows
    g = torch.Generator(device=dev).manual_seed(lvl * 100 + cin)
    x = torch.randn(n, cin, device=dev, generator=g)
    w = torch.randn(27, cin, cout, device=dev, generator=g) * 0.03
    wp = Fn.pack_weights(w, False, False, 3)
    Fn.bf16_planes(x, True)
    Fn.set_precision(0)
    ref = Fn.spconv_gather(x, table, w, cin, cout, False, False).double()
    Fn.set_precision(3)
    line = [f"L{lvl} n={n} {cin}->{cout}:"]
    for name, fuse, sets in variants:
        if fuse == 2 and 2 * cout > 256:
            continue
        raw.us3d_debug_set_tuning4(0, 0, 0, fuse)
        raw.us3d_debug_set_tuning_acc(sets)
        for with_bn in (False, True):
            ts = []
            for _ in range(7):
                flush.zero_()
                req = Fn.BnRequest(None, None, 0.1, 1e-5, None) if with_bn else None
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                y = Fn.spconv_gather(x, table, w, cin, cout, False, False, wpack=wp, bn=req)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            err = float((y.double() - ref).norm() / ref.norm())
            assert err < 1e-4, (name, err)
            line.append(f"{name}{'+bn' if with_bn else ''} {sorted(ts)[3]:.1f}")
    raw.us3d_debug_set_tuning4(0, 0, 0, 0)
    raw.us3d_debug_set_tuning_acc(0)
    print(" | ".join(line), flush=True)
