"""Runs the Mask3D forward on the CUDA stack and on the CPU oracle with forward hooks on every module and prints, in
execution order, the modules whose outputs deviate (diagnostics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import engine, models
from golden.make_golden import mask3d_inputs, MASK3D_KW
from helpers import Cfg, deterministic_state, our_models_on_oracle
from oracle import me_cpu


def tensors_of(out):
    if isinstance(out, torch.Tensor):
        return [out]
    if hasattr(out, "F") and hasattr(out, "coordinate_map_key"):
        return [out.F]
    if isinstance(out, (list, tuple)):
        r = []
        for o in out:
            r += tensors_of(o)
        return r
    if isinstance(out, dict):
        r = []
        for k in sorted(out):
            if k in ("pred_logits", "pred_masks"):
                r += tensors_of(out[k])
        return r
    return []


def run(models_pkg, me, device):
    coords, feats, raw, p2s, targets = mask3d_inputs()
    backbone = models_pkg.res16unet.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    net = models_pkg.mask3d.Mask3D(type("C", (), {"backbone": backbone})(), **MASK3D_KW)
    net.load_state_dict(deterministic_state(net, 7))
    net = net.to(device).train()
    log = []
    names = {m: n for n, m in net.named_modules()}

    def hook(mod, inp, out):
        log.append((names[mod], [t.detach().float().cpu() for t in tensors_of(out) if t.dtype != torch.bool] +
                    [t.detach().float().cpu() for t in tensors_of(out) if t.dtype == torch.bool]))

    for m in net.modules():
        m.register_forward_hook(hook)
    x = me.SparseTensor(feats.to(device), torch.from_numpy(coords).to(device))
    net(x, point2segment=[p.to(device) for p in p2s], raw_coordinates=raw.to(device))
    return log, coords


log_g, coords = run(models, engine, "cuda")
log_c, _ = run(our_models_on_oracle(), me_cpu, "cpu")
n0 = int((coords[:, 0] == 0).sum())
print("calls", len(log_g), len(log_c), "scene-0 voxels", n0, "of", coords.shape[0])
shown = 0
from collections import defaultdict
by_name = defaultdict(list)
for nc, tc in log_c:
    by_name[nc].append(tc)
seen = defaultdict(int)
for ng, tg in log_g:
    k = seen[ng]; seen[ng] += 1
    if k >= len(by_name[ng]):
        continue
    tc = by_name[ng][k]
    for a, b in zip(tg, tc):
        if a.shape != b.shape:
            print(ng, k, "SHAPE", a.shape, b.shape); shown += 1; continue
        scale = float(b.abs().max()) + 1e-12
        err = float((a - b).abs().max()) / scale
        if err > 2e-3:
            bad = (a - b).abs().reshape(a.shape[0], -1).max(1)[0] > 2e-3 * scale if a.ndim >= 1 and a.shape[0] > 1 else None
            where = "" if bad is None else f" first bad index along dim0 {int(torch.nonzero(bad)[0])} of {a.shape[0]}, count {int(bad.sum())}"
            print(f"{ng:55s} call {k} shape {tuple(a.shape)} err/max {err:.3e}{where}")
            shown += 1
    if shown > 30:
        break
