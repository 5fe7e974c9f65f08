"""What bounds k_spconv_mt on the 200k-voxel 96 -> 96 layer?  Same sparsity pattern, but every present neighbour points into a
window of W rows (so the gather hits L1/L2 trivially): if the time collapses the kernel is bound by the gather's memory
behaviour, if it stays the pipeline / MMA issue is the limit."""
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
table = cm.forward_table(k1, k1, (3, 3, 3))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cin = cout = 96
x = torch.randn(s.n, cin, device=dev)
w = torch.randn(27, cin, cout, device=dev) * 0.03


def med(t):
    fn = lambda: Fn.spconv_gather(x, t, w, cin, cout, False, False)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[2]


for order_on in (0, 32768):
    engine.set_row_ordering(order_on)
    base = engine.NeighbourTable(table.nbr, table.mask, table.n_rows, table.kvol)
    print(f"row order {'pattern' if order_on else 'natural'}: real table {med(base):.3f} ms", flush=True)
    for W in (128, 4096, 65536):
        fake = torch.where(table.nbr >= 0, table.nbr % W, table.nbr)
        t = engine.NeighbourTable(fake.contiguous(), table.mask, table.n_rows, table.kvol)
        print(f"    neighbours folded into {W} rows: {med(t):.3f} ms", flush=True)
    dense = torch.where(table.nbr >= 0, table.nbr, torch.zeros_like(table.nbr))
    t = engine.NeighbourTable(dense.contiguous(), table.mask, table.n_rows, table.kvol)
    print(f"    no absent neighbours (all 27 fetched, from row 0 where absent): {med(t):.3f} ms", flush=True)
