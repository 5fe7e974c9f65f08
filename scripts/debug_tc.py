"""Bring-up diagnostics for the tcgen05 sparse-conv kernel: compares it against the exact fp32 SIMT kernel
on (1) a plain GEMM (kvol = 1, identity table) with structured inputs, (2) k3 convolutions over a random
scene for the backbone's channel pairs, (3) times the big-layer shapes on a 200k-voxel scene."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import unscene3d_b200  # noqa: F401
from helpers import random_scene
from unscene3d_b200 import engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.engine.coords import NeighbourTable

dev = torch.device("cuda")
Fn._tc_kernel["fwd"] = os.environ.get("US3D_TC_KERNEL", "mt")
print("tensor-core forward kernel:", Fn._tc_kernel["fwd"], flush=True)


def run(x, table, w3, cin, cout, mode, transpose=False, flip=False):
    Fn.set_precision(mode)
    y = Fn.spconv_gather(x, table, w3, cin, cout, transpose, flip)
    torch.cuda.synchronize()
    return y


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


def gemm_test(n, cin, cout):
    table = NeighbourTable(torch.arange(n, dtype=torch.int32, device=dev)[None].contiguous(), None, n, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, cin, generator=g).to(dev)
    w = (torch.randn(1, cin, cout, generator=g) * 0.1).to(dev)
    ref = run(x, table, w, cin, cout, 0)
    for mode in (1, 3):
        y = run(x, table, w, cin, cout, mode)
        e = rel(y, ref)
        print(f"GEMM n={n} {cin}->{cout} mode {mode}: rel err {e:.3e}", flush=True)
        if e > (3e-2 if mode == 1 else 1e-4):
            # structured probe: x = one-hot on channel c for row r=c, w[c, n] = c + n/1000
            xs = torch.zeros(n, cin, device=dev)
            for r in range(min(n, cin)):
                xs[r, r] = 1.0
            ws = (torch.arange(cin, device=dev)[:, None] + torch.arange(cout, device=dev)[None] / 1000.0)[None].contiguous().float()
            ys = run(xs, table, ws, cin, cout, mode)
            print("  probe rows 0..3 (expected row r = r + col/1000):")
            print(ys[:4, :8].cpu().numpy())
            print("  probe rows 8,9,16,17,64,65:")
            print(ys[[8, 9, 16, 17, 64, 65], :8].cpu().numpy())
            return False
    return True


def conv_test():
    c = random_scene(5000, 3, batch=2, extent=30)
    x0 = engine.SparseTensor(torch.zeros(c.shape[0], 1, device=dev), torch.from_numpy(c).to(dev))
    cm, key = x0.coordinate_manager, x0.coordinate_map_key
    table = cm.forward_table(key, key, (3, 3, 3))
    g = torch.Generator().manual_seed(1)
    ok = True
    for cin, cout in [(32, 32), (64, 64), (96, 96), (128, 96), (256, 256), (384, 256), (32, 64), (192, 128)]:
        x = torch.randn(c.shape[0], cin, generator=g).to(dev)
        w = (torch.randn(27, cin, cout, generator=g) * (1.0 / (cin * 11) ** 0.5)).to(dev)
        ref = run(x, table, w, cin, cout, 0)
        reft = run(x[:, :cout].contiguous() if cout <= cin else torch.randn(c.shape[0], cout, generator=g).to(dev), table, w, cout, cin, 0, True, True)
        for mode in (1, 3):
            y = run(x, table, w, cin, cout, mode)
            xin = x[:, :cout].contiguous() if cout <= cin else None
            msg = f"conv k3 {cin}->{cout} mode {mode}: fwd rel err {rel(y, ref):.3e}"
            if xin is not None:
                yt = run(xin, table, w, cout, cin, mode, True, True)
                msg += f"  dgrad rel err {rel(yt, reft):.3e}"
            print(msg, flush=True)
            ok = ok and rel(y, ref) < (3e-2 if mode == 1 else 1e-4)
    return ok


def timing():
    from unscene3d_b200.synthetic import make_scene

    s = make_scene(200_000, seed=0, with_masks=False)
    c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
    x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
    cm, key = x0.coordinate_manager, x0.coordinate_map_key
    table = cm.forward_table(key, key, (3, 3, 3))
    for cin, cout in [(96, 96), (128, 96)]:
        x = torch.randn(s.n, cin, device=dev)
        w = torch.randn(27, cin, cout, device=dev) * 0.03
        for mode in (0, 1, 3):
            Fn.set_precision(mode)
            for _ in range(2):
                Fn.spconv_gather(x, table, w, cin, cout, False, False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                Fn.spconv_gather(x, table, w, cin, cout, False, False)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            pairs = int((table.nbr >= 0).sum())
            print(f"200k k3 {cin}->{cout} mode {mode}: {ms:.3f} ms  ({2 * pairs * cin * cout / ms / 1e9:.1f} TFLOP/s on real pairs, "
                  f"{2 * 27 * s.n * cin * cout / ms / 1e9:.1f} dense-equivalent)", flush=True)


if __name__ == "__main__":
    ok = gemm_test(256, 64, 32)
    ok = ok and gemm_test(1000, 128, 96)
    ok = ok and gemm_test(300, 32, 256)
    if ok:
        ok = conv_test()
    if ok:
        timing()
    print("RESULT", "OK" if ok else "FAILED")
