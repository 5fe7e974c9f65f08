"""ME.utils.sparse_quantize / sparse_collate on HOST inputs: the reference calls them in forked DataLoader
workers (datasets/utils.py:266-287, 403-432), so they must work without a CUDA context.  These tests run on the CPU-only build
container — nothing here can touch a device — and compare the libus3d host function with the oracle, bit-exact."""
import multiprocessing as mp

import numpy as np
import pytest
import torch

from oracle import me_cpu


def _cases():
    rng = np.random.default_rng(0)
    yield rng.uniform(-3, 3, size=(5000, 3)), 0.25
    yield rng.uniform(-1, 1, size=(20000, 3)), 0.02          # the reference's voxel size, many duplicates
    yield rng.integers(-40, 40, size=(3000, 3)).astype(np.float64), None
    yield np.zeros((0, 3)), 0.5                                # empty cloud
    yield np.array([[0.1, 0.2, 0.3]]), 0.5                     # single point
    yield rng.uniform(-1000, 1000, size=(4000, 3)), 0.01       # voxel indices up to +-1e5 (the oracle packs +-2^17 per axis)


@pytest.mark.parametrize("case", range(6))
def test_host_sparse_quantize_matches_oracle(case):
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import engine

    pts, q = list(_cases())[case]
    labels = np.random.default_rng(case).integers(0, 4, size=pts.shape[0])
    feats = np.random.default_rng(case + 10).normal(size=(pts.shape[0], 6)).astype(np.float32)
    got = engine.sparse_quantize(pts, features=feats, labels=labels, return_index=True, return_inverse=True, quantization_size=q)
    exp = me_cpu.sparse_quantize(pts, features=feats, labels=labels, return_index=True, return_inverse=True, quantization_size=q)
    assert len(got) == len(exp) == 5
    for g, e in zip(got, exp):
        assert np.array_equal(np.asarray(g), np.asarray(e))
    # maps only, torch flavour
    gm = engine.sparse_quantize(torch.from_numpy(pts), return_maps_only=True, return_inverse=True, quantization_size=q)
    em = me_cpu.sparse_quantize(torch.from_numpy(pts), return_maps_only=True, return_inverse=True, quantization_size=q)
    for g, e in zip(gm, em):
        assert torch.equal(g, e)


def _worker(queue):
    torch.set_num_threads(1)  # what torch.utils.data's worker loop does first (OpenMP pools do not survive a fork)
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import engine

    rng = np.random.default_rng(5)
    pts = rng.uniform(-2, 2, size=(3000, 3))
    c, idx = engine.sparse_quantize(pts, return_index=True, quantization_size=0.05)
    bc, f = engine.sparse_collate([c, c[:10]], [pts[idx].astype(np.float32), pts[idx][:10].astype(np.float32)])
    queue.put((c.shape, tuple(bc.shape), bool(torch.cuda.is_initialized())))


def test_quantize_and_collate_in_a_forked_worker_do_not_initialise_cuda():
    """What a DataLoader worker does (datasets/utils.py:403-432): voxelise + collate in a forked process."""
    import unscene3d_b200  # noqa: F401  (parent has the library loaded, as the training process has)

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(q,))
    p.start()
    shape, bshape, cuda_up = q.get(timeout=120)
    p.join(timeout=30)
    assert p.exitcode == 0
    assert shape[1] == 3 and bshape == (shape[0] + 10, 4) and not cuda_up
