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


if __name__ == "__main__" and "--mask3d" not in sys.argv:
    main()


# ------------------------------------------------------------------------------------------------- Mask3D step
MASK3D_KW = dict(hidden_dim=128, num_queries=100, num_heads=8, dim_feedforward=1024, sample_sizes=[200, 800, 3200, 12800, 51200],
                 shared_decoder=True, num_classes=3, num_decoders=3, dropout=0.0, pre_norm=False,
                 positional_encoding_type="fourier", non_parametric_queries=True, train_on_segments=True,
                 normalize_pos_enc=True, use_level_embed=False, scatter_type="mean", hlevels=[0, 1, 2, 3],
                 use_np_features=False, voxel_size=0.02, max_sample_size=False, random_queries=False, gauss_scale=1.0,
                 random_query_both=False, random_normal=False)  # conf/model/mask3d.yaml:5-34
LOSS_WEIGHTS = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0, "loss_noise_robust": 0.0}  # trainer/trainer.py:68-79


def mask3d_inputs(n=1500, batch=2, seed=404, n_seg=40, n_tgt=6):
    coords = random_scene(n, seed, batch=batch, extent=24)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(coords.shape[0], 3, generator=g)
    raw = torch.from_numpy(coords[:, 1:]).float() * 0.02 + torch.rand(coords.shape[0], 3, generator=g) * 0.01
    bidx = torch.from_numpy(coords[:, 0])
    p2s, targets = [], []
    for b in range(batch):
        nb = int((bidx == b).sum())
        seg = torch.randint(0, n_seg, (nb,), generator=g)
        seg[:n_seg] = torch.arange(n_seg)
        owner = torch.randint(0, n_tgt + 2, (n_seg,), generator=g)  # segments owned by target t (>= n_tgt: background)
        owner[:n_tgt] = torch.arange(n_tgt)
        seg_mask = torch.stack([owner == t for t in range(n_tgt)])
        targets.append({"labels": torch.ones(n_tgt, dtype=torch.long), "segment_mask": seg_mask,
                        "masks": seg_mask[:, seg], "point2segment": seg})
        p2s.append(seg)
    return coords, feats, raw, p2s, targets


def _steer_attention_masks(net, me, record=None, override=None, mismatches=None):
    """Wraps Mask3D.mask_module.  The decoder thresholds pooled mask logits into boolean attention masks
    (models/mask3d.py:437-446): a DISCRETE decision inside an otherwise continuous computation, and with randomly
    initialised weights many logits sit near zero.  `record` collects the masks of every round (golden generation);
    `override` replays recorded masks into the run under test after counting how many entries the run itself decided
    differently (`mismatches`), so that the continuous outputs of both runs stay comparable to round-off."""
    inner = net.mask_module
    state = {"k": 0}

    def mask_module(*args, **kwargs):
        out = inner(*args, **kwargs)
        if not (isinstance(out, tuple) and len(out) == 3):
            return out
        cls, masks, attn = out
        bits = attn.F
        k = state["k"]
        state["k"] += 1
        if record is not None:
            record.append(bits.detach().cpu().numpy().astype(bool))
        if override is not None:
            want = torch.from_numpy(override[k]).to(bits.device)
            assert want.shape == bits.shape, (k, want.shape, bits.shape)
            if mismatches is not None:
                mismatches.append((int((want != bits).sum()), bits.numel()))
            attn = me.SparseTensor(features=want, coordinate_manager=attn.coordinate_manager, coordinate_map_key=attn.coordinate_map_key)
        return cls, masks, attn

    net.mask_module = mask_module


def pack_attention(record):
    out = {"attn_rounds": np.asarray(len(record))}
    for k, b in enumerate(record):
        out[f"attn{k}"] = np.packbits(b.reshape(-1))
        out[f"attn{k}_shape"] = np.asarray(b.shape)
    return out


def unpack_attention(gold):
    return [np.unpackbits(gold[f"attn{k}"])[: int(np.prod(gold[f"attn{k}_shape"]))].reshape(tuple(gold[f"attn{k}_shape"])).astype(bool)
            for k in range(int(gold["attn_rounds"]))]


def run_mask3d_case(models_pkg, me, matcher, device="cpu", criterion_cls=None, attn_record=None, attn_override=None,
                    attn_mismatches=None, inputs=None):
    """Full self-training step (Mask3D forward, Hungarian matching, set criterion, backward).  `inputs` replaces the small
    fixture scene (coords [N, 4] int32 numpy, feats, raw coordinates, point2segment list, targets list)."""
    coords, feats, raw, p2s, targets = inputs if inputs is not None else mask3d_inputs()
    backbone = models_pkg.res16unet.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    net = models_pkg.mask3d.Mask3D(type("C", (), {"backbone": backbone})(), **MASK3D_KW)
    net.load_state_dict(deterministic_state(net, 7))
    net = net.to(device).train()
    if attn_record is not None or attn_override is not None:
        _steer_attention_masks(net, me, attn_record, attn_override, attn_mismatches)
    weight_dict = dict(LOSS_WEIGHTS)
    for i in range(len(MASK3D_KW["hlevels"]) * MASK3D_KW["num_decoders"]):
        weight_dict.update({f"{k}_{i}": v for k, v in LOSS_WEIGHTS.items()})
    criterion_cls = criterion_cls or models_pkg.criterion.SetCriterion
    crit = criterion_cls(num_classes=3, matcher=matcher, weight_dict=weight_dict, eos_coef=0.1, losses=["labels", "masks"],
                         num_points=-1, oversample_ratio=3.0, importance_sample_ratio=0.75, class_weights=-1).to(device)
    tg = [{k: v.to(device) for k, v in t.items()} for t in targets]
    x = me.SparseTensor(feats.to(device), torch.from_numpy(coords).to(device))
    out = net(x, point2segment=[p.to(device) for p in p2s], raw_coordinates=raw.to(device))
    losses = crit(out, tg, mask_type="segment_mask")
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    res = {"pred_logits": out["pred_logits"].detach().double().cpu().numpy(), "sampled_coords": np.asarray(out["sampled_coords"], dtype=np.float64),
           "total_loss": np.asarray(float(total))}
    for b, m in enumerate(out["pred_masks"]):
        res[f"pred_masks{b}"] = m.detach().double().cpu().numpy()
    for k in ("loss_ce", "loss_mask", "loss_dice", "loss_ce_5", "loss_mask_11", "loss_dice_0"):
        res["L:" + k] = np.asarray(float(losses[k]))
    idx = matcher({k: v for k, v in out.items() if k != "aux_outputs"}, tg, "segment_mask")
    for b, (i, j) in enumerate(idx):
        res[f"match{b}"] = np.stack([i.numpy(), j.numpy()])
    params = dict(net.named_parameters())
    missing = [k for k, v in params.items() if v.grad is None and not k.startswith("backbone.final")]
    assert not missing, f"parameters without gradient: {missing[:8]} ({len(missing)} of {len(params)})"
    for k in ("mask_features_head.kernel", "class_embed_head.weight", "cross_attention.0.2.multihead_attn.in_proj_weight",
              "query_projection.layers.0.weight", "backbone.block8.1.conv2.kernel", "backbone.conv0p1s1.kernel"):
        gr = params[k].grad.detach().double().cpu().reshape(-1)
        res["gnorm:" + k] = np.asarray(gr.norm().item())
    return res


def main_mask3d():
    from oracle import me_cpu

    ref = reference_models_on_oracle()
    matcher = ref.matcher.HungarianMatcher(cost_class=2.0, cost_mask=5.0, cost_dice=2.0, cost_noise_robust=0.0, num_points=-1)
    record = []
    res = run_mask3d_case(ref, me_cpu, matcher, attn_record=record)
    res.update(pack_attention(record))
    path = os.path.join(HERE, "mask3d_step.npz")
    np.savez_compressed(path, **res)
    print("mask3d_step", "loss", float(res["total_loss"]), {k: float(v) for k, v in res.items() if k.startswith("L:")},
          os.path.getsize(path), "bytes")


if __name__ == "__main__" and "--mask3d" in sys.argv:
    main_mask3d()
