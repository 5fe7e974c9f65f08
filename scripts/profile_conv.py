"""One big-layer sparse conv (200k voxels, k3, 128->96) through the production kernel — target for `ncu`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene

mode = int(os.environ.get("US3D_MODE", "3"))
dev = torch.device("cuda")
s = make_scene(200_000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
table = cm.forward_table(key, key, (3, 3, 3))
cin, cout = 128, 96
x = torch.randn(s.n, cin, device=dev)
dy = torch.randn(s.n, cout, device=dev)
w = torch.randn(27, cin, cout, device=dev) * 0.03
Fn.set_precision(mode)
for _ in range(3):
    y = Fn.spconv_gather(x, table, w, cin, cout, False, False)
    dw = Fn.spconv_wgrad(x, table, dy, cin, cout)
torch.cuda.synchronize()
print("done", float(y.abs().mean()), float(dw.abs().mean()))
