"""Bring-up diagnostics for the tcgen05 weight-gradient kernel against the exact fp32 SIMT kernel."""
import os
import sys

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


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


def run(x, table, dy, cin, cout, mode):
    Fn.set_precision(mode)
    dw = Fn.spconv_wgrad(x, table, dy, cin, cout)
    torch.cuda.synchronize()
    return dw


def dense_probe(n, cin, cout):
    """kvol = 1 identity table: dW = X^T dY."""
    table = NeighbourTable(torch.arange(n, dtype=torch.int32, device=dev)[None].contiguous(), None, n, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, cin, generator=g).to(dev)
    dy = torch.randn(n, cout, generator=g).to(dev)
    ref = x.double().T @ dy.double()
    ok = True
    for mode in (0, 1, 3):
        dw = run(x, table, dy, cin, cout, mode)[0]
        e = rel(dw, ref)
        print(f"dense wgrad n={n} {cin}x{cout} mode {mode}: rel err {e:.3e}", flush=True)
        if e > {0: 1e-5, 1: 3e-2, 3: 1e-4}[mode]:
            ok = False
            xs = torch.zeros(n, cin, device=dev)
            ds = torch.zeros(n, cout, device=dev)
            xs[:, :] = torch.arange(cin, device=dev)[None].float() + 1          # X[j, ci] = ci + 1
            ds[0, :] = torch.arange(cout, device=dev).float() / 100 + 1          # only row 0 contributes
            dws = run(xs, table, ds, cin, cout, mode)[0]
            print("  probe (expected dW[ci, co] = (ci+1) * (1 + co/100)):")
            print(dws[:4, :6].cpu().numpy())
            print(dws[[8, 9, 64, 65], :6].cpu().numpy() if cin > 65 else dws[[8, 9], :6].cpu().numpy())
    return ok


def conv_wgrad():
    c = random_scene(5000, 3, batch=2, extent=30)
    x0 = engine.SparseTensor(torch.zeros(c.shape[0], 1, device=dev), torch.from_numpy(c).to(dev))
    cm, key = x0.coordinate_manager, x0.coordinate_map_key
    table = cm.forward_table(key, key, (3, 3, 3))
    g = torch.Generator().manual_seed(1)
    ok = True
    for cin, cout in [(32, 32), (64, 64), (96, 96), (128, 96), (256, 256), (384, 256), (32, 64), (192, 128), (8, 16)]:
        x = torch.randn(c.shape[0], cin, generator=g).to(dev)
        dy = torch.randn(c.shape[0], cout, generator=g).to(dev)
        ref = run(x, table, dy, cin, cout, 0)
        for mode in (1, 3):
            dw = run(x, table, dy, cin, cout, mode)
            e = rel(dw, ref)
            print(f"k3 wgrad {cin}x{cout} mode {mode}: rel err {e:.3e}", flush=True)
            ok = ok and e < (3e-2 if mode == 1 else 1e-4)
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
        dy = torch.randn(s.n, cout, device=dev)
        for mode in (0, 1, 3):
            Fn.set_precision(mode)
            for _ in range(2):
                Fn.spconv_wgrad(x, table, dy, cin, cout)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                Fn.spconv_wgrad(x, table, dy, cin, cout)
            b.record()
            torch.cuda.synchronize()
            print(f"200k k3 wgrad {cin}x{cout} mode {mode}: {a.elapsed_time(b) / 5:.3f} ms", flush=True)


if __name__ == "__main__":
    ok = dense_probe(64, 128, 64)
    ok = dense_probe(1000, 128, 96) and ok
    ok = dense_probe(300, 32, 256) and ok
    if ok:
        ok = conv_wgrad()
    if ok:
        timing()
    print("RESULT", "OK" if ok else "FAILED")
