"""Timings of the other BASELINE.json configurations (not the headline metric; numbers quoted in DESIGN.md):

  C3  full Mask3D self-training step (Res16UNet34C backbone + mask decoder + Hungarian matcher + set criterion, forward +
      backward), batch of 4 synthetic 200k-voxel scenes with 20 pseudo masks each, 1 GPU
  C4  NCut pseudo masks: per-segment aggregation + affinity + greedy NCut extraction on a 300k-point scene with random
      384-d / 96-d features around per-segment centres
  C5  data-parallel self-training step, 2 scenes of 200k voxels per GPU, bucketed gradient all-reduce over NCCL
      (run under torchrun; with WORLD_SIZE=1 the all-reduce is skipped)

    python scripts/bench_configs.py [c3] [c4] [c5]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import numpy as np
import torch
import torch.distributed as dist

import unscene3d_b200  # noqa: F401
from unscene3d_b200 import distributed as D
from unscene3d_b200 import engine, models
from unscene3d_b200 import pseudo_masks as pm
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import collate, make_scene
from unscene3d_b200.utils import BackboneConfig, seeded_state

MASK3D_KW = dict(hidden_dim=128, num_queries=100, num_heads=8, dim_feedforward=1024, sample_sizes=[200, 800, 3200, 12800, 51200],
                 shared_decoder=True, num_classes=3, num_decoders=3, dropout=0.0, pre_norm=False,
                 positional_encoding_type="fourier", non_parametric_queries=True, train_on_segments=True,
                 normalize_pos_enc=True, use_level_embed=False, scatter_type="mean", hlevels=[0, 1, 2, 3],
                 use_np_features=False, voxel_size=0.02, max_sample_size=False, random_queries=False, gauss_scale=1.0,
                 random_query_both=False, random_normal=False)  # conf/model/mask3d.yaml:5-34
LOSS_WEIGHTS = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0, "loss_noise_robust": 0.0}


def build_mask3d(dev):
    backbone = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
    net = models.Mask3D(type("C", (), {"backbone": backbone})(), **MASK3D_KW)
    net.load_state_dict(seeded_state(net, 0))
    net = net.to(dev).train()
    weight_dict = dict(LOSS_WEIGHTS)
    for i in range(len(MASK3D_KW["hlevels"]) * MASK3D_KW["num_decoders"]):
        weight_dict.update({f"{k}_{i}": v for k, v in LOSS_WEIGHTS.items()})
    matcher = models.HungarianMatcher(cost_class=2.0, cost_mask=5.0, cost_dice=2.0, cost_noise_robust=0.0, num_points=-1)
    crit = models.SetCriterion(num_classes=3, matcher=matcher, weight_dict=weight_dict, eos_coef=0.1, losses=["labels", "masks"],
                               num_points=-1, oversample_ratio=3.0, importance_sample_ratio=0.75, class_weights=-1).to(dev)
    return net, crit, weight_dict


def scene_batch(n_scenes, n_voxels, seed0, dev):
    scenes = [make_scene(n_voxels, seed=seed0 + i, with_masks=True) for i in range(n_scenes)]
    coords, feats = collate(scenes)
    targets = [{"labels": torch.from_numpy(s.labels).to(dev), "segment_mask": torch.from_numpy(s.segment_mask).to(dev),
                "masks": torch.from_numpy(s.masks).to(dev), "point2segment": torch.from_numpy(s.point2segment).to(dev)} for s in scenes]
    return (torch.from_numpy(coords).to(dev), torch.from_numpy(feats[:, :3]).to(dev), torch.from_numpy(feats[:, 3:]).to(dev),
            [t["point2segment"] for t in targets], targets)


def train_step(net, crit, weight_dict, batch, world):
    coords, colors, raw, p2s, targets = batch
    Fn.pack_network(net)
    x = engine.SparseTensor(colors, coords)
    out = net(x, point2segment=p2s, raw_coordinates=raw)
    losses = crit(out, targets, mask_type="segment_mask")
    total = sum(losses[k] * weight_dict[k] for k in losses if k in weight_dict)
    total.backward()
    if world > 1:
        D.allreduce_gradients(net.parameters())
    net.zero_grad(set_to_none=True)
    return total


def timed_steps(fn, steps, warm):
    import gc

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    gc.collect()
    gc.freeze()  # long-lived objects out of the cyclic collector's way (a full pass over the heap is a ~100 ms host stall)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / steps


def run_c3(dev, rank, world, scenes_per_gpu=4, voxels=200_000, label="C3"):
    net, crit, wd = build_mask3d(dev)
    batch = scene_batch(scenes_per_gpu, voxels, 100 + rank * scenes_per_gpu, dev)
    ms = timed_steps(lambda: train_step(net, crit, wd, batch, world), steps=10, warm=4)
    ms = D.max_over_ranks(ms, dev)
    if rank == 0:
        n = scenes_per_gpu * world
        print(f"{label}: Mask3D self-train step, {scenes_per_gpu} x {voxels}-voxel scenes per GPU, {world} GPU(s): {ms:.1f} ms/step "
              f"-> {n / ms * 1e3:.2f} scenes/s, {n * voxels / ms * 1e3 / 1e6:.2f} M voxels/s", flush=True)


def run_c4(dev, n_points=300_000, feat_dims=(384, 96)):
    scene = make_scene(n_points, seed=7, with_masks=False)
    seg = torch.from_numpy(scene.point2segment).to(dev)
    S = scene.num_segments
    g = torch.Generator().manual_seed(0)
    n_clusters = 40
    owner = torch.randint(0, n_clusters, (S,), generator=g)
    feats = []
    for d in feat_dims:  # each segment's rows share a centre drawn around its cluster's direction, sigma_within = 0.3
        centres = torch.randn(n_clusters, d, generator=g)[owner] + 0.5 * torch.randn(S, d, generator=g)
        feats.append((centres[scene.point2segment] + 0.3 * torch.randn(n_points, d, generator=g)).to(dev))
    conn = torch.from_numpy(np.concatenate([scene.adjacency, scene.adjacency[:, ::-1]])).to(dev)
    torch.cuda.synchronize()

    def run():
        agg = tuple(pm.aggregate_features(f, seg, conn)[0] for f in feats)
        uniq = torch.unique(seg)
        return pm.unscene3d(agg, uniq, conn, affinity_tau=0.6, min_segment_size=4, max_extent_ratio=0.8, max_number_of_instances=20)

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    masks = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"C4: NCut pseudo masks, {n_points} points, S = {S} segments, features {feat_dims}: {dt:.2f} s per scene "
          f"({masks.shape[0]} masks; 20 NCut iterations incl. aggregation)", flush=True)


def main():
    which = [a.lower() for a in sys.argv[1:]] or ["c3", "c4", "c5"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    engine.set_coordinate_stream(torch.cuda.Stream(device=dev, priority=-1), dev)
    if "c3" in which and world == 1:
        run_c3(dev, rank, world)
    if "c4" in which and rank == 0 and world == 1:
        run_c4(dev)
    if "c5" in which:
        run_c3(dev, rank, world, scenes_per_gpu=2, label="C5")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
