"""Generates tests/golden/*.npz by running the UNMODIFIED reference model files
(/root/reference/models/res16unet.py, resnet.py, modules/*.py) on top of the CPU oracle
(oracle/me_cpu.py).  Only runs in the build container (the reference cannot travel):

    python tests/golden/make_golden.py

Inputs and weights are regenerated from seeds by the tests (helpers.random_scene,
helpers.deterministic_state), so the fixtures hold outputs only: per-level shapes, a strided sample
of every returned feature map, the scalar loss, and gradient samples of a few parameters.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import Cfg, deterministic_state, random_scene, reference_models_on_oracle  # noqa: E402

CASES = {
    # name: (class, n_voxels per scene, batch, scene seed, weight seed)
    "res16unet14": ("Res16UNet14", 2500, 2, 101, 1),
    "res16unet34c": ("Res16UNet34C", 3000, 2, 202, 2),
    "res16unet34c_multires": ("Res16UNet34CMultiRes", 2000, 1, 303, 3),
}
GRAD_KEYS = ["conv0p1s1.kernel", "block1.0.conv1.kernel", "block4.0.downsample.0.kernel", "convtr5p8s2.kernel",
             "block8.0.conv2.kernel", "bn0.bn.weight", "block8.0.norm1.bn.bias"]


def case_inputs(name):
    cls, n, batch, sseed, wseed = CASES[name]
    coords = random_scene(n, sseed, batch=batch, extent=28)
    g = torch.Generator().manual_seed(sseed)
    feats = torch.randn(coords.shape[0], 3, generator=g)
    return cls, coords, feats, wseed


def sample_rows(t: torch.Tensor, k: int = 48):
    idx = torch.linspace(0, t.shape[0] - 1, min(k, t.shape[0])).long()
    return t[idx].detach().double().numpy(), idx.numpy()


def run_case(models_pkg, me, name, device="cpu"):
    """Shared by the generator and the tests: returns dict of numpy results."""
    cls, coords, feats, wseed = case_inputs(name)
    net = getattr(models_pkg.res16unet, cls)(3, 20, Cfg(), D=3, out_fpn=True)
    net.load_state_dict(deterministic_state(net, wseed))
    net = net.to(device).train()
    x = me.SparseTensor(feats.to(device), torch.from_numpy(coords).to(device))
    out = net(x)
    if isinstance(out[1], dict):
        maps = [out[0]] + [out[1][k] for k in ("res_16", "res_8", "res_4", "res_2", "res_1")]
    else:
        maps = [out[0]] + list(out[1])
    res = {}
    loss = 0.0
    for i, m in enumerate(maps):
        # coarse maps are compared as sets: canonical order = sorted by (b, x, y, z)
        C = m.C.long().cpu()
        order = np.lexsort((C[:, 3].numpy(), C[:, 2].numpy(), C[:, 1].numpy(), C[:, 0].numpy()))
        F = m.F[torch.from_numpy(order).to(m.F.device)]
        res[f"shape{i}"] = np.asarray(F.shape)
        res[f"rows{i}"], res[f"idx{i}"] = sample_rows(F.cpu())
        res[f"sum{i}"] = np.asarray(F.detach().double().sum().item())
        res[f"abs{i}"] = np.asarray(F.detach().double().abs().sum().item())
        w = torch.linspace(-1, 1, F.shape[1], device=F.device)
        loss = loss + (F * w).mean() + (F * F).mean() * 0.1
    loss.backward()
    res["loss"] = np.asarray(loss.item())
    params = dict(net.named_parameters())
    for k in GRAD_KEYS:
        g = params[k].grad.detach().double().cpu().reshape(-1)
        res["grad:" + k] = g[:: max(g.numel() // 64, 1)][:64].numpy()
        res["gnorm:" + k] = np.asarray(g.norm().item())
    bn = dict(net.named_buffers())
    res["bn0.running_mean"] = bn["bn0.bn.running_mean"].detach().double().cpu().numpy()
    res["block1.0.norm1.running_var"] = bn["block1.0.norm1.bn.running_var"].detach().double().cpu().numpy()
    return res


def main():
    from oracle import me_cpu

    ref = reference_models_on_oracle()
    import models.res16unet  # noqa: F401  (the reference's, now in sys.modules)

    for name in CASES:
        res = run_case(ref, me_cpu, name)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **res)
        print(name, {k: v.tolist() for k, v in res.items() if k.startswith("shape")}, "loss", float(res["loss"]),
              os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
