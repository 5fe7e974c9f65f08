"""2D -> 3D feature lifting: us3d_project_features_2d3d behind the `project_features_cuda` module name.

GPU: (i) against the REFERENCE KERNEL ITSELF (oracle/_ref/libproject_ref.so = the reference's project_image_cuda_kernel.cu compiled
for sm_100a by oracle/build_ref.py): identical hit counts per voxel — both march with the same float expressions — and feature sums
at fp32 atomic round-off; (ii) the UNMODIFIED reference module utils/cuda_utils/raycast_image.py::Project2DFeaturesCUDA running on
the shims (`project_features_cuda`, `raycast_cuda`, `MinkowskiEngine.SparseTensor.dense`) against the numpy restatement
(oracle/ops_cpu.project_features_2d3d), which differs only for rays that graze a cell boundary within float round-off."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest
import torch

from helpers import staged_reference_root

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libproject_ref.so")


def room(n_voxels, seed):
    from unscene3d_b200.synthetic import make_scene

    s = make_scene(n_voxels, seed=seed, with_masks=False)
    c = s.coords.astype(np.int64)
    return c - c.min(0)


def camera(coords, seed, yaw_deg):
    """Camera-to-grid pose in voxel units: somewhere in the middle of the room at 1.5 m, looking horizontally."""
    rng = np.random.default_rng(seed)
    centre = coords.mean(0) + rng.normal(0, 5, 3)
    centre[2] = coords[:, 2].min() + 75
    yaw = np.deg2rad(yaw_deg)
    fwd = np.array([np.cos(yaw), np.sin(yaw), -0.15])
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0, 0, 1.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    m = np.eye(4, dtype=np.float32)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, down, fwd, centre  # camera x = right, y = down, z = forward
    return m


def dense_grid(coords):
    X, Y, Z = (coords.max(0) + 1).tolist()
    occ = np.zeros((1, Z, Y, X), dtype=np.int64)
    occ[0, coords[:, 2], coords[:, 1], coords[:, 0]] = np.arange(coords.shape[0])
    return occ


@pytest.mark.gpu
@pytest.mark.parametrize("n_voxels,W,H,C,yaw", [(20000, 64, 48, 8, 20.0), (60000, 160, 120, 384, 200.0), (60000, 320, 240, 16, 95.0)])
def test_cuda_projection_equals_the_reference_kernel(n_voxels, W, H, C, yaw):
    import unscene3d_b200  # noqa: F401
    import project_features_cuda as ours

    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libproject_ref.so not built (oracle/build_ref.py)")
    coords = room(n_voxels, 3)
    occ = torch.from_numpy(dense_grid(coords)).cuda()
    view = torch.from_numpy(camera(coords, 1, yaw)).view(1, 1, 4, 4).cuda()
    intr = torch.tensor([[0.9 * W, 0.9 * W, W / 2 - 0.5, H / 2 - 0.5]], dtype=torch.float32).cuda()
    feats = torch.randn(1, 1, H, W, C, generator=torch.Generator().manual_seed(0)).cuda()
    opts = torch.FloatTensor([W, H, 0.1 / 0.02, 4.0 / 0.02, 0.5])
    n = coords.shape[0]
    cnt_a, out_a = torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, C, device="cuda")
    ours.project_features_cuda(feats, occ, view, intr, opts, cnt_a, out_a, torch.BoolTensor([False]))
    cnt_b, out_b = torch.zeros(n, dtype=torch.int32, device="cuda"), torch.zeros(n, C, device="cuda")
    ref = ctypes.CDLL(REF_LIB)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    ref.project_ref.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 8 + [ctypes.c_float] * 3 + [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    torch.cuda.synchronize()
    Z, Y, X = occ.shape[1:]
    assert ref.project_ref(ptr(feats), ptr(occ), ptr(view), ptr(intr), 1, 1, H, W, C, Z, Y, X, float(opts[2]), float(opts[3]), float(opts[4]), n,
                           ptr(cnt_b), ptr(out_b)) == 0
    assert int(cnt_b.sum()) > 0.5 * H * W, "the camera sees too little of the room for a meaningful comparison"
    assert torch.equal(cnt_a, cnt_b), f"{int((cnt_a != cnt_b).sum())} voxels with different hit counts"
    assert float((out_a - out_b).abs().max()) <= 1e-5 * max(float(out_b.abs().max()), 1.0)
    assert int(cnt_a[0]) == 0  # occupancy value 0 = empty: voxel 0 is never hit


@pytest.mark.gpu
@pytest.mark.skipif(staged_reference_root() is None, reason="no reference tree (neither /root/reference nor oracle/_ref/reference)")
def test_reference_projection_module_on_the_shims_matches_the_oracle():
    import unscene3d_b200  # noqa: F401
    from oracle import ops_cpu

    path = os.path.join(staged_reference_root(), "utils", "cuda_utils", "raycast_image.py")
    spec = importlib.util.spec_from_file_location("ref_raycast_image", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # its `import MinkowskiEngine / raycast_cuda / project_features_cuda` resolve to the shims
    coords = room(20000, 5)
    shift = np.array([7, -3, 11])
    batched = torch.from_numpy(np.concatenate([np.zeros((coords.shape[0], 1), np.int64), coords + shift], 1)).int().cuda()
    W, H, C = 48, 36, 12
    pose = camera(coords, 2, 140.0)
    pose[:3, 3] += shift
    view = torch.from_numpy(pose).view(1, 1, 4, 4).cuda()
    intr = torch.tensor([[0.8 * W, 0.8 * W, W / 2 - 0.5, H / 2 - 0.5]], dtype=torch.float32).cuda()
    feats = torch.randn(1, 1, H, W, C, generator=torch.Generator().manual_seed(1)).cuda()
    cfg = type("Cfg", (), {"data": type("D", (), {"ignore_label": 255})()})()
    proj = mod.Project2DFeaturesCUDA(width=W, height=H, voxel_size=0.02, config=cfg)
    got, counts = proj(feats, batched, view, intr)
    hit, want_counts, sums = ops_cpu.project_features_2d3d(feats.cpu().numpy(), dense_grid(coords), camera(coords, 2, 140.0)[None, None], intr.cpu().numpy(),
                                                           proj.depth_min, proj.depth_max, proj.ray_increment)
    counts = counts.cpu().numpy()
    assert want_counts.sum() > 0.5 * H * W
    differs = counts != want_counts
    assert differs.sum() <= max(4, 2e-3 * (want_counts > 0).sum()), f"{differs.sum()} voxels with different hit counts"
    same = ~differs & (want_counts > 0)
    want = sums[same] / (want_counts[same, None] + 10e-5)
    assert np.abs(got.cpu().numpy()[same] - want).max() <= 1e-5 * max(np.abs(want).max(), 1.0)
