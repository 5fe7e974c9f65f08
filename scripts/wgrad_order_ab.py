"""Weight gradient on the 200k-voxel k3 map: natural table vs pattern-ordered table with permuted dY planes."""
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
k2 = cm.stride(k1, (2, 2, 2))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for key, cin, cout in ((k1, 96, 96), (k1, 128, 96), (k2, 96, 96), (k2, 32, 32)):
    table = cm.forward_table(key, key, (3, 3, 3))
    x = torch.randn(table.n_rows, cin, device=dev)
    dy = torch.randn(table.n_rows, cout, device=dev)
    res = {}
    for mode in ("natural", "permute"):
        Fn._wgrad_order["mode"] = mode
        for _ in range(2):
            dw = Fn.spconv_wgrad(x, table, dy, cin, cout)
        ts = []
        for _ in range(5):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dw = Fn.spconv_wgrad(x, table, dy, cin, cout)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[mode] = (sorted(ts)[2], dw)
    err = float((res["permute"][1] - res["natural"][1]).norm() / res["natural"][1].norm())
    print(f"n={table.n_rows} {cin}->{cout}: natural {res['natural'][0]:.3f} ms, permute {res['permute'][0]:.3f} ms (incl. permute pass + zero fill), rel diff {err:.2e}", flush=True)
