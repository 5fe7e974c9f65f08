"""Tri-plane projection of the noise-robust loss (SURVEY §8(b) `custom_cuda_utils`).

CPU: the numpy restatement (oracle/ops_cpu.project_voxels_to_planes[_bwd], following cuda_utils_kernel.cu:371-433, 496-556) on a
hand-checked case.  GPU: us3d_project_voxels_to_planes[_bwd] through the `custom_cuda_utils` shim against (i) the restatement,
(ii) the REFERENCE KERNELS themselves (oracle/_ref/libplanes_ref.so = the reference's cuda_utils_kernel.cu compiled for sm_100a by
oracle/build_ref.py), and (iii) the unmodified reference module models/noise_robust_loss.py running on the shim, forward and
backward, against this repo's module and a dense torch-autograd restatement.  Sums are fp32 atomics in both implementations:
1e-5 relative; counts and the skip rule are exact.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import staged_reference_root

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(REPO, "oracle", "_ref", "libplanes_ref.so")


def make_case(n, inst, seed, extent=(14, 11, 9)):
    rng = np.random.default_rng(seed)
    c = np.unique(np.stack([rng.integers(0, e + 1, n) for e in extent], 1), axis=0)
    c = c[rng.permutation(c.shape[0])]
    coords = np.concatenate([np.zeros((c.shape[0], 1), np.int64), c], 1).astype(np.int32)
    coords[:, 1:] += np.array([-3, 5, 2], np.int32)  # the wrapper centres the coordinates itself
    pred = rng.random((c.shape[0], inst)).astype(np.float32)
    tgt = (rng.random((c.shape[0], inst)) < 0.3).astype(np.float32)
    return coords, pred, tgt


def test_oracle_projection_on_a_hand_checked_case():
    from oracle import ops_cpu

    coords = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 1, 0, 0], [0, 2, 1, 0], [0, 1, 1, 2]])  # dims = max = (2, 1, 2)
    pred = np.array([[1.0], [2.0], [4.0], [8.0], [16.0]])
    out = ops_cpu.project_voxels_to_planes(coords, pred, pred * 0, (2, 1, 2))
    # voxels 3 (x = 2 = x_dim, y = 1 = y_dim) and 4 (y = 1, z = 2) are outside: only voxels 0, 1, 2 count
    assert out["xy"][2].tolist() == [[2], [1]] and out["xy"][0][:, :, 0].tolist() == [[3.0], [4.0]]
    assert out["xz"][2].tolist() == [[1, 1], [1, 0]] and out["yz"][0][:, :, 0].tolist() == [[5.0, 2.0]]
    g = {"xy": np.array([[[3.0]], [[0.0]]]), "xz": np.array([[[1.0], [0.0]], [[5.0], [0.0]]]), "yz": np.array([[[2.0], [0.0]]])}
    back = ops_cpu.project_voxels_to_planes_bwd(coords, g, (2, 1, 2), 1)
    assert np.allclose(back[:, 0], [(3 + 1 + 2) / 3, 3.0 / 1, (5 + 2) / 2, 0.0, 0.0])


def _planes(dims, inst, dev):
    x, y, z = dims
    shapes = ((x, y), (x, z), (y, z))
    return ([torch.zeros((*s, inst), device=dev) for s in shapes], [torch.zeros((*s, inst), device=dev) for s in shapes],
            [torch.zeros(s, device=dev, dtype=torch.int32) for s in shapes])


@pytest.mark.gpu
@pytest.mark.parametrize("n,inst,seed", [(1, 1, 0), (300, 7, 1), (5000, 20, 2), (200000, 13, 3)])
def test_cuda_projection_matches_oracle_and_reference_kernels(n, inst, seed):
    import unscene3d_b200  # noqa: F401
    import custom_cuda_utils
    from oracle import ops_cpu

    extent = (14, 11, 9) if n < 100000 else (120, 90, 60)
    coords, pred, tgt = make_case(n, inst, seed, extent)
    c = coords.copy()
    c[:, 1:] -= c[:, 1:].min(0)
    dims = tuple(int(v) for v in c[:, 1:].max(0))
    dev = torch.device("cuda")
    cd, pd, td = torch.from_numpy(c).to(dev), torch.from_numpy(pred).to(dev), torch.from_numpy(tgt).to(dev)
    P, T, N = _planes(dims, inst, dev)
    custom_cuda_utils.project_sparse_voxels_to_planes(cd, pd, td, *P, *T, *N)
    want = ops_cpu.project_voxels_to_planes(c, pred, tgt, dims)
    for k, name in enumerate(("xy", "xz", "yz")):
        assert np.array_equal(N[k].cpu().numpy(), want[name][2]), name
        for got, w in ((P[k], want[name][0]), (T[k], want[name][1])):
            assert w.size == 0 or np.abs(got.cpu().numpy() - w).max() <= 1e-5 * max(np.abs(w).max(), 1.0), name
    g = [torch.randn_like(p) * (torch.rand_like(p) < 0.8) for p in P]  # some exact zeros: the backward counts non-zero views
    back = torch.full((c.shape[0], inst), 7.0, device=dev)
    custom_cuda_utils.project_sparse_voxels_to_planes_backward(cd, back, *g, *N)
    wb = ops_cpu.project_voxels_to_planes_bwd(c, {k: v.cpu().numpy() for k, v in zip(("xy", "xz", "yz"), g)}, dims, inst)
    inside = (c[:, 1] < dims[0]) & (c[:, 2] < dims[1]) & (c[:, 3] < dims[2])
    assert not inside.any() or np.abs(back.cpu().numpy()[inside] - wb[inside]).max() <= 1e-6
    assert (back.cpu().numpy()[~inside] == 7.0).all(), "skipped voxels must be left untouched, as the reference leaves them"
    if os.path.exists(REF_LIB) and min(dims) > 0:
        ref = ctypes.CDLL(REF_LIB)
        P2, T2, N2 = _planes(dims, inst, dev)
        torch.cuda.synchronize()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        assert ref.planes_ref_fwd(ptr(cd), ptr(pd), ptr(td), c.shape[0], inst, *dims, *[ptr(t) for t in P2 + T2 + N2]) == 0
        for a, b in zip(P + T, P2 + T2):
            assert float((a - b).abs().max()) <= 1e-5 * max(float(b.abs().max()), 1.0)
        for a, b in zip(N, N2):
            assert torch.equal(a, b)
        back2 = torch.full((c.shape[0], inst), 7.0, device=dev)
        torch.cuda.synchronize()
        assert ref.planes_ref_bwd(ptr(cd), ptr(back2), c.shape[0], inst, *dims, *[ptr(t) for t in g + N2]) == 0
        assert torch.equal(back, back2), "backward differs from the reference kernel"


def _dense_restatement(logits, targets, coords, directions="xyz"):
    """ProjectionMaskLoss as plain differentiable torch ops (index_add on flattened planes) — autograd then yields
    d loss / d logits through the MEANS, which equals the reference's hand-written backward only where every voxel's three
    cell gradients are non-zero; used for the forward value."""
    c = (coords - coords.amin(0))[:, 1:].long()
    dims = c.max(0)[0]
    ok = (c < dims).all(1)
    p = torch.sigmoid(logits.T)[ok]
    t = targets.T[ok]
    c = c[ok]
    loss, cells = 0.0, 0
    for axis, (a, b) in (("z", (0, 1)), ("y", (0, 2)), ("x", (1, 2))):
        cell = c[:, a] * dims[b] + c[:, b]
        n_cells = int(dims[a] * dims[b])
        num = torch.zeros(n_cells, device=p.device).index_add_(0, cell, torch.ones_like(cell, dtype=torch.float32))
        ps = torch.zeros(n_cells, p.shape[1], device=p.device).index_add_(0, cell, p) / (num[:, None] + 10e-9)
        ts = torch.zeros(n_cells, p.shape[1], device=p.device).index_add_(0, cell, t) / (num[:, None] + 10e-9)
        occ = num > 0
        cells += int(occ.sum())
        if axis in directions:
            loss = loss + torch.nn.functional.binary_cross_entropy(ps[occ], ts[occ], reduction="sum")
    return loss, logits.shape[0] * cells


@pytest.mark.gpu
def test_noise_robust_loss_modules_on_the_shim():
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.models.noise_robust_loss import ProjectionMaskLoss

    coords, pred, tgt = make_case(4000, 6, 9)
    dev = torch.device("cuda")
    cd = torch.from_numpy(coords).to(dev)
    logits = torch.from_numpy(np.log(pred / (1 - pred + 1e-6) + 1e-6).T.copy()).to(dev).requires_grad_()
    targets = torch.from_numpy(tgt.T.copy()).to(dev)
    ours = ProjectionMaskLoss(directions="xyz")
    loss, shape = ours(logits, targets, cd)
    loss.backward()
    g_ours = logits.grad.clone()
    want, want_shape = _dense_restatement(logits.detach(), targets, cd)
    assert shape == want_shape and abs(float(loss) - float(want)) <= 1e-4 * abs(float(want))
    root = staged_reference_root()
    if root is not None:  # the reference's own module file, unmodified, over the same shim
        from helpers import reference_models_on_shim

        ref_models = reference_models_on_shim()
        import importlib

        nrl = importlib.import_module("models.noise_robust_loss")
        assert root in nrl.__file__
        logits2 = logits.detach().clone().requires_grad_()
        loss2, shape2 = nrl.ProjectionMaskLoss(directions="xyz")(logits2, targets, cd)
        loss2.backward()
        assert shape2 == shape and abs(float(loss2) - float(loss)) <= 1e-5 * abs(float(loss))
        assert float((logits2.grad - g_ours).abs().max()) <= 1e-5 * float(g_ours.abs().max())
