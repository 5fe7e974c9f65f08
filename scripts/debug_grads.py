"""Debug helper: per-parameter gradient error of the CUDA backbone vs the CPU oracle (fp32 and fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
import unscene3d_b200
from helpers import Cfg, deterministic_state, our_models_on_oracle, random_scene
from oracle import me_cpu
from unscene3d_b200 import engine, models

arch = sys.argv[1] if len(sys.argv) > 1 else "Res16UNet14"
coords = random_scene(4000, 5, batch=2, extent=30)
feats = torch.randn(coords.shape[0], 3, generator=torch.Generator().manual_seed(5))
cpu_net = getattr(our_models_on_oracle().res16unet, arch)(3, 20, Cfg(), D=3, out_fpn=True)
state = deterministic_state(cpu_net, 3)
cpu_net.load_state_dict(state)
cpu_net = cpu_net.double()
gpu_net = getattr(models, arch)(3, 20, Cfg(), D=3, out_fpn=True)
gpu_net.load_state_dict(state)
gpu_net.cuda()
fg = feats.cuda().requires_grad_()
fc = feats.double().requires_grad_()
out_g, maps_g = gpu_net(engine.SparseTensor(fg, torch.from_numpy(coords).cuda()))
out_c, maps_c = cpu_net(me_cpu.SparseTensor(fc, torch.from_numpy(coords)))
C = out_c.F.shape[1]
w = torch.linspace(-1, 1, C)
(out_g.F * w.cuda()).mean().backward()
(out_c.F * w.double()).mean().backward()
rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp(min=1e-300))
print("features", rel(out_g.F, out_c.F), "input grad", rel(fg.grad, fc.grad))
pc, pg = dict(cpu_net.named_parameters()), dict(gpu_net.named_parameters())
for k in pc:
    if pc[k].grad is None: continue
    print(f"{k:40s} {tuple(pc[k].shape)!s:20s} {rel(pg[k].grad, pc[k].grad):.3e}")
