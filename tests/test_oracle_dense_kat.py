"""Pins the CPU oracle (oracle/me_cpu.py) with known answers that do not depend on anybody's memory
of MinkowskiEngine: on a fully occupied block every sparse operator must equal its dense torch
counterpart (SURVEY.md §8(c) KATs i–v), plus float64 gradcheck on random sparse scenes.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import me_cpu as ME


def dense_block(B, X, Y, Z, origin=(0, 0, 0)):
    g = np.stack(np.meshgrid(np.arange(B), np.arange(X), np.arange(Y), np.arange(Z), indexing="ij"), -1).reshape(-1, 4)
    g[:, 1:] += np.asarray(origin)
    return g.astype(np.int32)


def to_dense(st: ME.SparseTensor, origin, shape, stride=1):
    C = st.C.long()
    B = int(C[:, 0].max()) + 1
    out = torch.zeros((B, st.F.shape[1], *shape), dtype=st.F.dtype)
    idx = (C[:, 1:] - torch.tensor(origin)) // stride
    out[C[:, 0], :, idx[:, 0], idx[:, 1], idx[:, 2]] = st.F
    return out


def torch_weight(kernel, ks):
    """W_torch[co, ci, ix, iy, iz] = W[ix + ks*iy + ks^2*iz, ci, co]  (x fastest, Appendix A.4)."""
    K, ci, co = kernel.shape
    return kernel.reshape(ks, ks, ks, ci, co).permute(4, 3, 2, 1, 0).contiguous()


@pytest.mark.parametrize("origin", [(0, 0, 0), (-4, -6, -2)])
def test_k3s1_equals_conv3d(origin):
    torch.manual_seed(0)
    B, X, Y, Z, ci, co = 2, 6, 5, 4, 3, 5
    c = dense_block(B, X, Y, Z, origin)
    perm = torch.randperm(c.shape[0]).numpy()
    c = c[perm]
    f = torch.randn(c.shape[0], ci, dtype=torch.float64)
    conv = ME.MinkowskiConvolution(ci, co, kernel_size=3, stride=1, dimension=3).double()
    y = conv(ME.SparseTensor(f, torch.from_numpy(c)))
    xd = to_dense(ME.SparseTensor(f, torch.from_numpy(c)), origin, (X, Y, Z))
    yd = F.conv3d(xd, torch_weight(conv.kernel.detach(), 3), padding=1)
    assert torch.allclose(to_dense(y, origin, (X, Y, Z)), yd, atol=1e-12)


@pytest.mark.parametrize("origin", [(0, 0, 0), (-4, -6, -2)])
def test_k2s2_equals_strided_conv3d(origin):
    torch.manual_seed(1)
    B, X, Y, Z, ci, co = 2, 6, 4, 8, 4, 3
    c = dense_block(B, X, Y, Z, origin)
    f = torch.randn(c.shape[0], ci, dtype=torch.float64)
    conv = ME.MinkowskiConvolution(ci, co, kernel_size=2, stride=2, dimension=3).double()
    x = ME.SparseTensor(f, torch.from_numpy(c))
    y = conv(x)
    assert y.tensor_stride == [2, 2, 2]
    assert y.F.shape[0] == B * X * Y * Z // 8
    yd = F.conv3d(to_dense(x, origin, (X, Y, Z)), torch_weight(conv.kernel.detach(), 2), stride=2)
    assert torch.allclose(to_dense(y, origin, (X // 2, Y // 2, Z // 2), stride=2), yd, atol=1e-12)


def test_k2s2_transpose_equals_conv_transpose3d():
    torch.manual_seed(2)
    B, X, Y, Z, ci, co = 1, 4, 6, 4, 3, 2
    origin = (-2, 0, -4)
    c = dense_block(B, X, Y, Z, origin)
    fine = ME.SparseTensor(torch.zeros(c.shape[0], 1, dtype=torch.float64), torch.from_numpy(c))
    down = ME.MinkowskiConvolution(1, ci, kernel_size=2, stride=2, dimension=3).double()
    coarse = down(fine)
    coarse = ME.SparseTensor(torch.randn(coarse.F.shape[0], ci, dtype=torch.float64),
                             coordinate_map_key=coarse.coordinate_map_key, coordinate_manager=coarse.coordinate_manager)
    up = ME.MinkowskiConvolutionTranspose(ci, co, kernel_size=2, stride=2, dimension=3).double()
    y = up(coarse)
    assert y.coordinate_map_key == fine.coordinate_map_key  # lands on the existing finer map (A.2/A.7)
    # conv_transpose3d weight layout [ci, co, kx, ky, kz]
    w = up.kernel.detach().reshape(2, 2, 2, ci, co).permute(3, 4, 2, 1, 0).contiguous()
    yd = F.conv_transpose3d(to_dense(coarse, origin, (X // 2, Y // 2, Z // 2), stride=2), w, stride=2)
    assert torch.allclose(to_dense(y, origin, (X, Y, Z)), yd, atol=1e-12)


def test_avg_pool_equals_avg_pool3d_and_counts_present_children_only():
    torch.manual_seed(3)
    B, X, Y, Z, ch = 2, 4, 4, 6, 5
    c = dense_block(B, X, Y, Z)
    f = torch.randn(c.shape[0], ch, dtype=torch.float64)
    pool = ME.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)
    x = ME.SparseTensor(f, torch.from_numpy(c))
    y = pool(x)
    yd = F.avg_pool3d(to_dense(x, (0, 0, 0), (X, Y, Z)), 2)
    assert torch.allclose(to_dense(y, (0, 0, 0), (X // 2, Y // 2, Z // 2), stride=2), yd, atol=1e-12)
    # drop voxels: mean is over the present ones
    keep = torch.rand(c.shape[0]) < 0.4
    keep[0] = True
    xs = ME.SparseTensor(f[keep], torch.from_numpy(c[keep.numpy()]))
    ys = pool(xs)
    Cs = xs.C.long()
    parent = torch.cat([Cs[:, :1], Cs[:, 1:] // 2 * 2], 1)
    for r in range(ys.F.shape[0]):
        m = (parent == ys.C[r].long()).all(1)
        assert torch.allclose(ys.F[r], xs.F[m].mean(0), atol=1e-12)


def test_negative_coordinates_floor_division():
    c = torch.tensor([[0, -1, -2, -3], [0, -4, 0, 1], [0, 3, -1, 0]], dtype=torch.int32)
    x = ME.SparseTensor(torch.ones(3, 1), c)
    y = ME.MinkowskiSumPooling(kernel_size=2, stride=2, dimension=3)(x)
    got = {tuple(r) for r in y.C.tolist()}
    assert got == {(0, -2, -2, -4), (0, -4, 0, 0), (0, 2, -2, 0)}


def test_1x1_conv_is_mm_with_2d_kernel_and_bias():
    torch.manual_seed(4)
    conv = ME.MinkowskiConvolution(6, 4, kernel_size=1, stride=1, bias=True, dimension=3)
    assert conv.kernel.shape == (6, 4) and conv.bias.shape == (1, 4)
    c = torch.from_numpy(dense_block(1, 3, 3, 3))
    f = torch.randn(27, 6)
    y = conv(ME.SparseTensor(f, c))
    assert torch.allclose(y.F, f @ conv.kernel + conv.bias)


def test_first_occurrence_order_and_quantize_maps():
    pts = np.array([[0.5, 0.5, 0.5], [1.2, 0.1, 0.3], [0.9, 0.2, 0.1], [-0.1, 0.0, 0.0], [1.7, 0.9, 0.2]])
    uc, um, im = ME.sparse_quantize(pts, return_index=True, return_inverse=True)
    assert um.tolist() == [0, 1, 3]
    assert im.tolist() == [0, 1, 0, 2, 1]
    assert uc.tolist() == [[0, 0, 0], [1, 0, 0], [-1, 0, 0]]
    bc, bf = ME.sparse_collate([torch.from_numpy(uc), torch.from_numpy(uc[:2])], [torch.ones(3, 2), torch.zeros(2, 2)])
    assert bc.dtype == torch.int32 and bc[:, 0].tolist() == [0, 0, 0, 1, 1] and bf.shape == (5, 2)


def test_gradcheck_sparse_conv_chain():
    torch.manual_seed(5)
    rng = np.random.default_rng(0)
    c = np.unique(rng.integers(-5, 5, size=(200, 3)), axis=0)
    c = np.concatenate([np.zeros((c.shape[0], 1), np.int64), c], 1).astype(np.int32)
    c3 = ME.MinkowskiConvolution(2, 3, kernel_size=3, stride=1, dimension=3).double()
    dn = ME.MinkowskiConvolution(3, 3, kernel_size=2, stride=2, dimension=3).double()
    up = ME.MinkowskiConvolutionTranspose(3, 2, kernel_size=2, stride=2, dimension=3).double()

    def fn(f, w3, wd, wu):
        x = ME.SparseTensor(f, torch.from_numpy(c))
        cm = x.coordinate_manager
        k3 = cm.kernel_map(x.coordinate_map_key, x.coordinate_map_key, (3, 3, 3))
        y = ME.sparse_conv_forward(f, w3, k3, f.shape[0])
        ck = cm.stride(x.coordinate_map_key, (2, 2, 2))
        kd = cm.kernel_map(x.coordinate_map_key, ck, (2, 2, 2))
        z = ME.sparse_conv_forward(y, wd, kd, cm.size(ck))
        return ME.sparse_conv_forward(z, wu, kd, f.shape[0], transpose_map=True)

    f = torch.randn(c.shape[0], 2, dtype=torch.float64, requires_grad=True)
    args = (f, c3.kernel.detach().requires_grad_(), dn.kernel.detach().requires_grad_(), up.kernel.detach().requires_grad_())
    assert torch.autograd.gradcheck(fn, args, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# SURVEY §8(c) KATs (vi) and (vii): the two host-side solvers the path leans on
def test_hungarian_assignment_is_the_brute_force_optimum():
    """(vi) Cost matrices of the matcher's shape ([Q queries, T <= 6 targets], oracle cost function): the assignment
    scipy.optimize.linear_sum_assignment returns — the solver models/matcher.py:161-163 and our matcher both call — has the
    cost of the best of all Q!/(Q-T)! injections."""
    import itertools

    from scipy.optimize import linear_sum_assignment

    from oracle import ops_cpu

    g = torch.Generator().manual_seed(0)
    for T in range(1, 7):
        Q, S = 7, 40
        logits = torch.randn(Q, 3, generator=g)
        masks = torch.randn(S, Q, generator=g) * 2
        tgt = torch.rand(T, S, generator=g) < 0.3
        labels = torch.randint(0, 2, (T,), generator=g)
        cost = np.asarray(ops_cpu.matcher_cost(logits, masks, tgt, labels, 2.0, 5.0, 2.0), dtype=np.float64)
        assert cost.shape == (Q, T)
        i, j = linear_sum_assignment(cost)
        best = min(sum(cost[q, t] for t, q in enumerate(perm)) for perm in itertools.permutations(range(Q), T))
        assert abs(cost[i, j].sum() - best) < 1e-9
        assert sorted(j.tolist()) == list(range(T)) and len(set(i.tolist())) == T


def test_fiedler_vector_is_the_second_eigenvector_of_the_normalised_laplacian():
    """(vii) oracle.ncut_cpu.fiedler (the reference's scipy eigh(D - A, D, subset_by_index=[1, 2]),
    pseudo_masks/unscene3d_pseudo_main.py:138-146) against numpy.linalg.eigh of D^-1/2 (D - A) D^-1/2: v = D^-1/2 u_2 up to
    sign and scale, on a thresholded {1, eps} affinity of two loosely coupled blobs."""
    from oracle import ncut_cpu

    rng = np.random.default_rng(1)
    n = 60
    A = np.full((n, n), 1e-5)
    A[:35, :35] = 1.0
    A[35:, 35:] = 1.0
    flip = rng.random((n, n)) < 0.03
    flip = np.triu(flip, 1)
    flip = flip | flip.T
    A[flip] = np.where(A[flip] == 1.0, 1e-5, 1.0)
    np.fill_diagonal(A, 1.0)
    D = np.diag(A.sum(1))
    v = np.asarray(ncut_cpu.fiedler(A, D)).reshape(-1)
    d = np.diag(D)
    L = (D - A) / np.sqrt(np.outer(d, d))
    w, U = np.linalg.eigh(L)
    assert w[0] < 1e-10 < w[1] < w[2] - 1e-6          # simple second eigenvalue
    ref = U[:, 1] / np.sqrt(d)
    cos = abs(v @ ref) / (np.linalg.norm(v) * np.linalg.norm(ref))
    assert cos > 1 - 1e-10
    # and it separates the two blobs
    side = v > v.mean()
    assert side[:35].all() != side[35:].all() and (side[:35].all() or (~side[:35]).all())
