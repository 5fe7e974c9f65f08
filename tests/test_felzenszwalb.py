"""Felzenszwalb mesh over-segmentation: the `felzenszwalb_cpp` shim over libus3d's host function against the
REFERENCE MODULE ITSELF — utils/cpp_utils/segmentator.cpp compiled unmodified into oracle/_ref/felzenszwalb_ref by
oracle/build_ref.py.  Segment ids and the adjacency list must be identical arrays (integer work: bit-exact), on meshes with many
equal edge weights (flat uniformly coloured regions), degenerate faces (NaN normals) and small segments that get merged."""
import importlib.util
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "oracle", "_ref", "felzenszwalb_ref")


def reference_module():
    import glob

    paths = glob.glob(os.path.join(REF_DIR, "felzenszwalb_cpp*.so"))
    if not paths:
        return None
    spec = importlib.util.spec_from_file_location("felzenszwalb_cpp", paths[0])  # the module's init symbol needs this name
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def grid_mesh(nx, ny, seed, bumps=6, palette=5, degenerate=0):
    """Height-field mesh with a few bumps, piecewise-constant colours plus noise on some patches, optional zero-area faces."""
    rng = np.random.default_rng(seed)
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float32), np.arange(ny, dtype=np.float32), indexing="ij")
    z = np.zeros_like(xs)
    for _ in range(bumps):
        cx, cy, r, h = rng.uniform(0, nx), rng.uniform(0, ny), rng.uniform(2, 6), rng.uniform(1, 4)
        z += np.where((xs - cx) ** 2 + (ys - cy) ** 2 < r * r, h, 0).astype(np.float32)
    z += (1.5 * np.sin(xs * 0.45) * np.cos(ys * 0.3)).astype(np.float32) + rng.normal(0, 0.15, xs.shape).astype(np.float32)
    v = np.stack([xs, ys, z], -1).reshape(-1, 3).astype(np.float32) * np.float32(0.05)
    cell = (xs // 7).astype(np.int64) * 131 + (ys // 5).astype(np.int64)
    colours = rng.random((int(cell.max()) + 1, 3)).astype(np.float32)
    colours = colours[rng.integers(0, palette, colours.shape[0])] if palette else colours
    c = colours[cell].reshape(-1, 3).copy()
    noisy = rng.random(c.shape[0]) < 0.3
    c[noisy] += (rng.normal(0, 0.02, (int(noisy.sum()), 3))).astype(np.float32)
    idx = np.arange(nx * ny).reshape(nx, ny)
    a, b, cc, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, 1:].ravel()
    f = np.concatenate([np.stack([a, b, cc], 1), np.stack([b, d, cc], 1)]).astype(np.int32)
    f = f[rng.permutation(f.shape[0])]
    if degenerate:
        f[:degenerate, 2] = f[:degenerate, 1]  # zero-area faces: the face normal is 0 / 0
    return v, f, c.astype(np.float32)


@pytest.mark.skipif(reference_module() is None, reason="oracle/_ref/felzenszwalb_ref not built (oracle/build_ref.py)")
@pytest.mark.parametrize("nx,ny,seed,min_verts,degenerate", [(12, 9, 0, 4, 0), (60, 45, 1, 20, 0), (150, 120, 2, 20, 0), (80, 80, 3, 50, 7),
                                                             (200, 160, 4, 20, 3)])
def test_segment_mesh_equals_the_reference_module(nx, ny, seed, min_verts, degenerate):
    import unscene3d_b200  # noqa: F401
    import felzenszwalb_cpp as ours

    assert "shims" in ours.__file__
    ref = reference_module()
    v, f, c = grid_mesh(nx, ny, seed, degenerate=degenerate)
    want_comps, want_conn = ref.segment_mesh(v, f, c, 0.005, min_verts)
    got_comps, got_conn = ours.segment_mesh(v, f, c, 0.005, min_verts)
    assert np.array_equal(got_comps, np.asarray(want_comps)), "segment ids differ"
    assert np.array_equal(got_conn, np.asarray(want_conn).reshape(-1, 2)), "segment adjacency differs"
    n_seg = int(got_comps.max()) + 1
    assert n_seg >= 2 and (np.bincount(got_comps) > 0).all()
    # no segment below the minimum size survives next to a neighbour
    sizes = np.bincount(got_comps, minlength=n_seg)
    small = np.nonzero(sizes < min_verts)[0]
    assert not np.isin(got_conn[:, 0], small).any() if got_conn.size else True


def test_segment_mesh_properties_without_the_reference():
    """Connected uniformly coloured flat mesh -> one segment; two flat sheets of very different colour joined at a crease ->
    the crease separates them; ids are compact and the adjacency is symmetric-free of self pairs."""
    import unscene3d_b200  # noqa: F401
    import felzenszwalb_cpp as ours

    v, f, c = grid_mesh(20, 20, 0, bumps=0, palette=0)
    c[:] = 0.5
    comps, conn = ours.segment_mesh(v, f, c, 0.005, 20)
    assert comps.max() == 0 and conn.shape == (0, 2)
    v2 = v.copy()
    right = v2[:, 0] > v2[:, 0].mean()
    v2[right, 2] = (v2[right, 0] - v2[:, 0].mean()) * 3.0  # a steep ramp
    c2 = c.copy()
    c2[right] = np.array([1.0, 0.0, 0.0], np.float32)
    comps, conn = ours.segment_mesh(v2, f, c2, 0.005, 20)
    assert len(np.unique(comps)) >= 2 and len(np.unique(comps[~right & (v2[:, 0] < v2[:, 0].mean() - 0.1)])) == 1
    assert (conn[:, 0] != conn[:, 1]).all() and np.array_equal(np.unique(comps), np.arange(comps.max() + 1))
