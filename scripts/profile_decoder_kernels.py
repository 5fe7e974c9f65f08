"""FPS (4 scenes x 200k voxels x 100 samples) and masked cross-attention (C3 shape of hlevel 3: B=4, K=12800, Q=100, h=8) — targets
for `ncu`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import unscene3d_b200  # noqa
from unscene3d_b200.engine import functional as Fn

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
pts = torch.randint(-300, 300, (4, 200_000, 3), device=dev, generator=g).float()
Q, K, B, H, hd = 100, 12800, 4, 8, 16
q = torch.randn(Q, B, H * hd, device=dev, generator=g, requires_grad=True)
k = torch.randn(K, B, H * hd, device=dev, generator=g, requires_grad=True)
v = torch.randn(K, B, H * hd, device=dev, generator=g, requires_grad=True)
mask = Fn.DecoderMask(torch.rand(B, K, Q, device=dev, generator=g) < 0.7)
for _ in range(3):
    idx = Fn.furthest_point_sampling(pts, 100)
    out = Fn.MaskedCrossAttentionFunction.apply(q, k, v, mask, H)
    out.sum().backward()
torch.cuda.synchronize()
print("done", int(idx.sum()), float(out.abs().mean()))
