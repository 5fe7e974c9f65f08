"""FreeMask-style pseudo-mask variant (pseudo_masks/freemask_main.py:203-417).

CPU: the oracle (oracle/freemask_cpu.py) against tests/golden/freemask_scene.npz, which was produced by executing the
reference's own source lines (tests/golden/make_freemask_golden.py) — bit-exact.
GPU: unscene3d_b200.pseudo_masks.freemask (libus3d kernels + host set logic) against the golden and against the oracle on
further seeded scenes: the selected masks must be the same point sets in the same order, soft values and maskness within 1e-5.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import freemask_cpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "freemask_scene.npz")


def _gold_case():
    g = np.load(GOLD)
    return {k[3:]: g[k] for k in g.files if k.startswith("in_")}, g["soft_masks"], g["maskness"]


def _oracle(case, trace=None):
    return freemask_cpu.freemask(torch.from_numpy(case["keys_F"]), torch.from_numpy(case["matching_segment_ids"]),
                                 torch.from_numpy(case["seg_connectivity"]), case["lr_coords"], torch.from_numpy(case["coords"]), trace=trace)


def test_oracle_reproduces_reference_golden_bit_exact():
    case, soft, maskness = _gold_case()
    trace = {}
    got_soft, got_maskness = _oracle(case, trace)
    assert np.array_equal(got_soft.numpy(), soft)
    assert np.array_equal(got_maskness.numpy(), maskness)
    # the case exercises the separation step: more blobs than queries
    assert trace["separated_segments"].shape[0] > trace["soft_segments"].shape[0]


def test_separation_keeps_the_reference_index_skip():
    """A bridging segment that touches three earlier blobs merges only two of them: after `pop` the reference advances its
    index past the blob that slid into the freed slot (freemask_main.py:306-318)."""
    unique = torch.arange(7)
    # blobs {0}, {2}, {4} exist when segment 5 (adjacent to 0, 2 and 4) arrives
    conn = {0: set(), 1: set(), 2: set(), 3: set(), 4: set(), 5: {0, 2, 4}, 6: set()}
    masks = torch.tensor([[1, 0, 1, 0, 1, 1, 0]], dtype=torch.bool)
    blobs = freemask_cpu.separate_blobs(masks, unique, conn)[0]
    assert sorted(sorted(int(x) for x in b) for b in blobs) == [[0, 2, 5], [4]]


def _product(case):
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import pseudo_masks as pm

    return pm.freemask(torch.from_numpy(case["keys_F"]).cuda(), torch.from_numpy(case["matching_segment_ids"]).cuda(),
                       torch.from_numpy(case["seg_connectivity"]).cuda(), case["lr_coords"], torch.from_numpy(case["coords"]).cuda())


def _same_masks(got, ref, thr=freemask_cpu.DEFAULTS.hard_mask_threshold):
    got_soft, got_maskness = got
    ref_soft, ref_maskness = ref
    assert got_soft.shape == ref_soft.shape
    assert torch.equal(got_soft >= thr, ref_soft >= thr)                 # same point sets, same order
    assert float((got_soft - ref_soft).abs().max()) < 1e-5
    assert float((got_maskness - ref_maskness).abs().max()) < 1e-5


@pytest.mark.gpu
def test_cuda_freemask_matches_reference_golden():
    case, soft, maskness = _gold_case()
    _same_masks(_product(case), (torch.from_numpy(soft), torch.from_numpy(maskness)))


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(side=10, n_objects=5, n_prototypes=5, seed=1), dict(side=18, n_objects=12, n_prototypes=5, seed=2, noise=0.7),
                                dict(side=24, n_objects=20, n_prototypes=6, seed=4, points_per_segment=9, dim=96)])
def test_cuda_freemask_matches_oracle(kw):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_freemask_golden import make_case

    case = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in make_case(**kw).items()}
    ref = _oracle(case)
    got = _product(case)
    if ref is None:
        assert got is None
    else:
        _same_masks(got, ref)


@pytest.mark.gpu
def test_cuda_freemask_skips_the_scene_like_the_reference():
    """Where the reference `continue`s (no candidate survives a filter) both implementations return None."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_freemask_golden import make_case

    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import pseudo_masks as pm

    case = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in make_case(side=6, n_objects=3, n_prototypes=3, seed=9).items()}
    strict = dict(nms_maskness_threshold=1.5)  # maskness is <= 1: nothing passes the final filter (freemask_main.py:412-415)
    cfg = type(freemask_cpu.DEFAULTS)(**{**vars(freemask_cpu.DEFAULTS), **strict})
    ref = freemask_cpu.freemask(torch.from_numpy(case["keys_F"]), torch.from_numpy(case["matching_segment_ids"]),
                                torch.from_numpy(case["seg_connectivity"]), case["lr_coords"], torch.from_numpy(case["coords"]), cfg=cfg)
    got = pm.freemask(torch.from_numpy(case["keys_F"]).cuda(), torch.from_numpy(case["matching_segment_ids"]).cuda(),
                      torch.from_numpy(case["seg_connectivity"]).cuda(), case["lr_coords"], torch.from_numpy(case["coords"]).cuda(), **strict)
    assert ref is None and got is None


def test_host_separation_matches_oracle_on_random_graphs():
    """The host set logic of libus3d (a host function of the C ABI: runs without a device) against the oracle's literal
    restatement, including multi-blob merges."""
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200._lib import lib

    rng = np.random.default_rng(0)
    for trial in range(20):
        S = int(rng.integers(5, 60))
        M = int(rng.integers(1, 12))
        edges = rng.integers(0, S, size=(int(rng.integers(S, 4 * S)), 2))
        masks = rng.random((M, S)) < 0.5
        conn = {i: set(int(b) for a, b in edges if a == i) for i in range(S)}
        ref = freemask_cpu.separate_blobs(torch.from_numpy(masks), torch.arange(S), conn)
        order = np.argsort(edges[:, 0], kind="stable")
        adj = edges[order, 1].astype(np.int32)
        adj_ptr = np.zeros(S + 1, dtype=np.int32)
        np.cumsum(np.bincount(edges[:, 0], minlength=S), out=adj_ptr[1:])
        m8 = np.ascontiguousarray(masks.astype(np.uint8))
        cap = int(m8.sum()) + 1
        bq, bp, bm = np.empty(cap, np.int32), np.empty(cap + 1, np.int32), np.empty(cap, np.int32)
        nb = lib.us3d_freemask_separate_h(m8.ctypes.data, M, S, adj_ptr.ctypes.data, adj.ctypes.data, bq.ctypes.data, bp.ctypes.data,
                                          bm.ctypes.data, cap, cap)
        got = [[] for _ in range(M)]
        for b in range(nb):
            got[bq[b]].append(sorted(int(x) for x in bm[bp[b]:bp[b + 1]]))
        want = [[sorted(int(x) for x in blob) for blob in q] for q in ref]
        assert got == want, trial
