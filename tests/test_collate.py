"""Voxelise + collate + target building.

CPU: the UNMODIFIED reference function datasets/utils.py::freemask_voxelize runs over the shim's host path
(`ME.utils.sparse_quantize / sparse_collate` on libus3d's host hash, no CUDA context) and over the oracle's restatement of
MinkowskiEngine — identical voxels, maps and targets.  GPU: unscene3d_b200.collate.freemask_voxelize_device (pinned staging, device
quantisation and target building) against the reference function, element for element."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

from helpers import minkowski_as, oracle_me_modules, staged_reference_root

needs_reference = pytest.mark.skipif(staged_reference_root() is None, reason="no reference tree (neither /root/reference nor oracle/_ref/reference)")


def load_reference_datasets_utils(me_modules=None, alias="ref_datasets_utils"):
    """datasets/utils.py imports only MinkowskiEngine, numpy, torch: loaded by path under `alias` with the given ME modules
    (None: whatever `import MinkowskiEngine` resolves to — the shim once unscene3d_b200 is imported)."""
    path = os.path.join(staged_reference_root(), "datasets", "utils.py")
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    if me_modules is None:
        spec.loader.exec_module(mod)
    else:
        with minkowski_as(me_modules):
            spec.loader.exec_module(mod)
    return mod


def make_batch(n_scenes=3, points=6000, seed=0, voxel_size=0.02):
    """Samples in the dataset's layout: several points per voxel (so quantisation merges), 2 + M freemask columns
    (label, masks, segment id) with a different M per scene (padding path) and one all-zero mask column."""
    rng = np.random.default_rng(seed)
    batch = []
    for s in range(n_scenes):
        n = points + 500 * s
        base = rng.uniform(-1.0, 1.0, size=(n // 3, 3))
        xyz = np.concatenate([base + rng.normal(0, 0.004, base.shape) for _ in range(3)])[:n]
        feats = rng.normal(size=(xyz.shape[0], 6)).astype(np.float32)
        seg = (np.floor((xyz[:, 0] + 1) * 4) * 64 + np.floor((xyz[:, 1] + 1) * 4) * 8 + 7 * s).astype(np.int64)
        m = 4 + s
        masks = np.zeros((xyz.shape[0], m), np.int64)
        for t in range(m - 1):  # instance t = the points of a few segments; the last column stays empty
            chosen = rng.choice(np.unique(seg), size=3, replace=False)
            masks[np.isin(seg, chosen), t] = 1
        fm = np.concatenate([np.zeros((xyz.shape[0], 1), np.int64), masks, seg[:, None]], 1)
        batch.append((xyz, feats, fm, f"scene{s:04d}", feats[:, :3].copy(), feats[:, 3:].copy(), xyz.copy(), s, np.zeros((0, 2), np.int64)))
    return batch


def compare_targets(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            assert torch.equal(g[k].cpu(), w[k].cpu()), k


@needs_reference
def test_reference_freemask_voxelize_on_the_shim_host_path_equals_the_oracle():
    import unscene3d_b200  # noqa: F401

    ref_shim = load_reference_datasets_utils(None, "ref_datasets_utils_shim")
    ref_oracle = load_reference_datasets_utils(oracle_me_modules(), "ref_datasets_utils_oracle")
    a = ref_shim.freemask_voxelize(make_batch(), 255, 0.02, "train", 100)
    b = ref_oracle.freemask_voxelize(make_batch(), 255, 0.02, "train", 100)
    assert torch.equal(a[0].coordinates, b[0].coordinates) and torch.equal(a[0].features, b[0].features)
    for x, y in zip(a[0].inverse_maps, b[0].inverse_maps):
        assert torch.equal(torch.as_tensor(x), torch.as_tensor(y))
    compare_targets(a[1], b[1])
    compare_targets(a[0].target_full, b[0].target_full)
    assert a[1][0]["segment_mask"].shape[0] == 3 and a[1][2]["masks"].shape[0] == 5  # the empty column is dropped


@pytest.mark.gpu
@needs_reference
@pytest.mark.parametrize("points,seed", [(6000, 0), (150000, 1)])
def test_device_voxelize_collate_targets_equal_the_reference_function(points, seed):
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.collate import freemask_voxelize_device

    ref = load_reference_datasets_utils(None, "ref_datasets_utils_shim")
    batch = make_batch(3, points, seed)
    want = ref.freemask_voxelize(batch, 255, 0.02, "train", 100)
    got = freemask_voxelize_device(batch, 0.02, "cuda")
    assert got["coordinates"].is_cuda and got["coordinates"].dtype == torch.int32
    assert torch.equal(got["coordinates"].cpu(), want[0].coordinates) and torch.equal(got["features"].cpu(), want[0].features)
    for x, y in zip(got["inverse_maps"], want[0].inverse_maps):
        assert torch.equal(x.cpu(), torch.as_tensor(y))
    compare_targets(got["target"], want[1])
    compare_targets(got["target_full"], want[0].target_full)
