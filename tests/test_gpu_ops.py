"""GPU parity tests, operator level: every libus3d kernel (through the C ABI) against the CPU oracle on
the same seeded inputs.  Integer results (coordinate maps, kernel maps, FPS indices) must be
bit-exact; floating point within the tolerance written next to each assert (north_star: 1e-3
relative on features; the exact-fp32 SIMT path is held to 1e-5).
"""
import numpy as np
import pytest
import torch

from helpers import random_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import unscene3d_b200  # noqa: F401  (loads libus3d.so, fails loudly if missing)
    from unscene3d_b200 import engine

    return engine


@pytest.fixture(scope="module")
def ora():
    from oracle import me_cpu

    return me_cpu


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


# --------------------------------------------------------------------------------------------- coords
@pytest.mark.parametrize("n,batch,seed", [(1, 1, 0), (37, 1, 1), (2000, 2, 2), (30000, 3, 3)])
def test_coordinate_maps_and_strides_bit_exact(eng, ora, n, batch, seed):
    c = random_scene(n, seed, batch=batch, extent=40)
    x = eng.SparseTensor(torch.zeros(c.shape[0], 1, device="cuda"), torch.from_numpy(c).cuda())
    y = ora.SparseTensor(torch.zeros(c.shape[0], 1), torch.from_numpy(c))
    assert torch.equal(x.C.cpu(), y.C)
    kx, ky = x.coordinate_map_key, y.coordinate_map_key
    for _ in range(4):
        kx = x.coordinate_manager.stride(kx, (2, 2, 2))
        ky = y.coordinate_manager.stride(ky, (2, 2, 2))
        assert kx.tensor_stride == ky.tensor_stride
        # same rows in the same (first-occurrence) order
        assert torch.equal(x.coordinate_manager.get_coordinates(kx).cpu(), y.coordinate_manager.get_coordinates(ky))


def test_duplicate_rows_keep_first_occurrence(eng, ora):
    rng = np.random.default_rng(7)
    base = random_scene(500, 5, batch=2, extent=12)
    dup = base[rng.integers(0, base.shape[0], size=1500)]
    feats = torch.arange(dup.shape[0], dtype=torch.float32)[:, None]
    x = eng.SparseTensor(feats.cuda(), torch.from_numpy(dup).cuda())
    y = ora.SparseTensor(feats, torch.from_numpy(dup))
    assert torch.equal(x.C.cpu(), y.C)
    assert torch.equal(x.F.cpu(), y.F)


def test_sparse_quantize_matches_oracle(eng, ora):
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 3, size=(5000, 3))
    labels = rng.integers(0, 4, size=5000)
    got = eng.sparse_quantize(pts, labels=labels, return_index=True, return_inverse=True, quantization_size=0.25)
    exp = ora.sparse_quantize(pts, labels=labels, return_index=True, return_inverse=True, quantization_size=0.25)
    for g, e in zip(got, exp):
        assert np.array_equal(np.asarray(g), np.asarray(e))
    # CUDA tensors go through the hash kernels (host inputs through the host function, tests/test_quantize.py)
    got = eng.sparse_quantize(torch.from_numpy(pts).cuda(), labels=torch.from_numpy(labels).cuda(), return_index=True, return_inverse=True,
                              quantization_size=0.25)
    for g, e in zip(got, exp):
        assert g.is_cuda and np.array_equal(g.cpu().numpy(), np.asarray(e))


def table_to_pairs(nbr):
    nbr = nbr.cpu().numpy()
    out = []
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        out.append(set(zip(nbr[k][o].tolist(), o.tolist())))
    return out


@pytest.mark.parametrize("ksize,stride", [((3, 3, 3), 1), ((2, 2, 2), 2), ((3, 3, 3), 2)])
def test_kernel_maps_bit_exact(eng, ora, ksize, stride):
    c = random_scene(6000, 11, batch=2, extent=30)
    x = eng.SparseTensor(torch.zeros(c.shape[0], 1, device="cuda"), torch.from_numpy(c).cuda())
    y = ora.SparseTensor(torch.zeros(c.shape[0], 1), torch.from_numpy(c))
    for level in range(3):
        kx_in, ky_in = x.coordinate_map_key, y.coordinate_map_key
        for _ in range(level):
            kx_in = x.coordinate_manager.stride(kx_in, (2, 2, 2))
            ky_in = y.coordinate_manager.stride(ky_in, (2, 2, 2))
        kx_out = x.coordinate_manager.stride(kx_in, (stride,) * 3)
        ky_out = y.coordinate_manager.stride(ky_in, (stride,) * 3)
        fwd = x.coordinate_manager.forward_table(kx_in, kx_out, ksize)
        exp = y.coordinate_manager.kernel_map(ky_in, ky_out, ksize)
        got = table_to_pairs(fwd.nbr)
        assert len(got) == len(exp)
        for k in range(len(exp)):
            assert got[k] == set(zip(exp[k][0].tolist(), exp[k][1].tolist())), f"offset {k} differs at level {level}"
        # transposed table = same pairs with roles swapped
        bwd, flip = x.coordinate_manager.backward_table(kx_in, kx_out, ksize)
        gotT = table_to_pairs(bwd.nbr)
        K = len(exp)
        for k in range(K):
            kk = K - 1 - k if flip else k
            assert gotT[kk] == {(o, i) for (i, o) in got[k]}
        # tile mask: bit k set iff the 128-row tile has a neighbour at offset k
        nbr = fwd.nbr.cpu().numpy()
        mask = fwd.mask.cpu().numpy().astype(np.uint32)
        for t in range(mask.shape[0]):
            want = 0
            for k in range(K):
                if (nbr[k, t * 128:(t + 1) * 128] >= 0).any():
                    want |= 1 << k
            assert int(mask[t]) == want


# --------------------------------------------------------------------------------------------- convs
def _pair(eng, ora, c, cin, seed=0):
    torch.manual_seed(seed)
    f = torch.randn(c.shape[0], cin)
    fx = f.clone().cuda().requires_grad_()
    fy = f.clone().requires_grad_()
    x = eng.SparseTensor(fx, torch.from_numpy(c).cuda())
    y = ora.SparseTensor(fy, torch.from_numpy(c))
    return x, y, fx, fy


def _sync_params(mx, my):
    my.load_state_dict({k: v.cpu() for k, v in mx.state_dict().items()})


@pytest.mark.parametrize("cin,cout,ks,stride,bias", [
    (3, 32, 3, 1, False), (32, 32, 2, 2, False), (32, 64, 3, 1, False), (64, 64, 3, 1, False), (128, 96, 3, 1, False),
    (96, 96, 3, 1, False), (256, 256, 3, 1, False), (384, 256, 1, 1, False), (96, 20, 1, 1, True), (5, 7, 3, 1, True),
    (16, 24, 3, 2, False), (384, 256, 3, 1, False), (3, 64, 2, 2, False), (3, 96, 1, 1, True),
])
@pytest.mark.parametrize("mode", [0, 3, 1])
def test_convolution_forward_backward(eng, ora, cin, cout, ks, stride, bias, mode):
    """mode 0 = exact fp32 SIMT kernel, 3 = tcgen05 three-term bf16 split (fp32-faithful), 1 = tcgen05 plain bf16.
    Shapes the tensor-core kernel does not take (cin or cout not a multiple of 16) fall to the SIMT kernel."""
    from unscene3d_b200.engine import functional as Fn

    tol = {0: 1e-5, 3: 5e-5, 1: 2e-2}[mode]
    Fn.set_precision(mode)
    try:
        _conv_fwd_bwd(eng, ora, cin, cout, ks, stride, bias, tol)
    finally:
        Fn.set_precision(3)


def _conv_fwd_bwd(eng, ora, cin, cout, ks, stride, bias, tol):
    c = random_scene(3000, 21, batch=2, extent=26)
    x, y, fx, fy = _pair(eng, ora, c, cin)
    mx = eng.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, bias=bias, dimension=3).cuda()
    my = ora.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, bias=bias, dimension=3)
    _sync_params(mx, my)
    ox, oy = mx(x), my(y)
    assert torch.equal(ox.C.cpu(), oy.C)
    assert rel_err(ox.F, oy.F) < tol
    g = torch.randn_like(oy.F)
    ox.F.backward(g.cuda())
    oy.F.backward(g)
    assert rel_err(fx.grad, fy.grad) < tol
    assert rel_err(mx.kernel.grad, my.kernel.grad) < tol
    if bias:
        assert rel_err(mx.bias.grad, my.bias.grad) < 1e-5


@pytest.mark.parametrize("cin,cout", [(256, 256), (128, 96), (96, 96), (6, 10)])
def test_transposed_convolution_forward_backward(eng, ora, cin, cout):
    """runs in the default precision (tcgen05 three-term split where the shape allows)"""
    c = random_scene(3000, 23, batch=2, extent=26)
    x, y, _, _ = _pair(eng, ora, c, 4)
    dx = eng.MinkowskiConvolution(4, cin, kernel_size=2, stride=2, dimension=3).cuda()
    dy = ora.MinkowskiConvolution(4, cin, kernel_size=2, stride=2, dimension=3)
    _sync_params(dx, dy)
    ux = eng.MinkowskiConvolutionTranspose(cin, cout, kernel_size=2, stride=2, dimension=3).cuda()
    uy = ora.MinkowskiConvolutionTranspose(cin, cout, kernel_size=2, stride=2, dimension=3)
    _sync_params(ux, uy)
    hx, hy = dx(x), dy(y)
    hx = eng.SparseTensor(hx.F.detach().requires_grad_(), coordinate_map_key=hx.coordinate_map_key, coordinate_manager=hx.coordinate_manager)
    hy = ora.SparseTensor(hy.F.detach().requires_grad_(), coordinate_map_key=hy.coordinate_map_key, coordinate_manager=hy.coordinate_manager)
    ox, oy = ux(hx), uy(hy)
    assert ox.coordinate_map_key.tensor_stride == (1, 1, 1)
    assert rel_err(ox.F, oy.F) < 5e-5
    g = torch.randn_like(oy.F)
    ox.F.backward(g.cuda())
    oy.F.backward(g)
    assert rel_err(hx.F.grad, hy.F.grad) < 5e-5
    assert rel_err(ux.kernel.grad, uy.kernel.grad) < 5e-5


def test_pattern_ordered_table_is_a_permutation_of_the_table(eng):
    """Integer work, bit-exact: order is a permutation (stable within equal patterns), column j of the re-ordered table is
    column order[j] of the table, tile masks are the unions over 128-row tiles, keys sort the rarest offset first."""
    from unscene3d_b200.engine import coords as C

    c = random_scene(5000, 5, batch=2, extent=30)
    x = eng.SparseTensor(torch.zeros(c.shape[0], 1, device="cuda"), torch.from_numpy(c).cuda())
    cm, key = x.coordinate_manager, x.coordinate_map_key
    C.set_row_ordering(1)
    try:
        table = cm.forward_table(key, key, (3, 3, 3))
        nbr_o, mask_o, order = (t.cpu().numpy() for t in table.ordered())
    finally:
        C.set_row_ordering(32768)
    nbr = table.nbr.cpu().numpy()
    n = nbr.shape[1]
    assert np.array_equal(np.sort(order), np.arange(n))
    assert np.array_equal(nbr_o, nbr[:, order])
    present = nbr >= 0
    freq = present.sum(1)
    pos = np.argsort(np.argsort(-freq, kind="stable"), kind="stable")      # most frequent offset -> bit 0, ties by offset index
    keys = (present.astype(np.int64) << pos[:, None]).sum(0)
    assert np.array_equal(order, np.argsort(keys, kind="stable"))
    tiles = (n + 127) // 128
    want = np.zeros(tiles, dtype=np.int64)
    for k in range(27):
        hit = np.zeros(tiles * 128, bool)
        hit[:n] = nbr_o[k] >= 0
        want |= hit.reshape(tiles, 128).any(1).astype(np.int64) << k
    assert np.array_equal(mask_o.astype(np.int64) & 0x7FFFFFF, want)
    # the point of the exercise: fewer active (tile, offset) pairs than in input order
    assert sum(bin(int(v)).count("1") for v in want) < sum(bin(int(v) & 0x7FFFFFF).count("1") for v in table.mask.cpu().numpy())


@pytest.mark.parametrize("cin,cout,ks,stride", [(96, 96, 3, 1), (32, 64, 3, 1), (32, 32, 2, 2), (128, 96, 3, 2)])
def test_convolution_on_pattern_ordered_tables(eng, ora, cin, cout, ks, stride):
    """Rows grouped by neighbour pattern (large maps in production; forced here): same results as the unordered run (per
    row the same products in the same order, only all-zero tile x offset products are skipped; the weight gradient sums
    rows in another order), and everything within tolerance of the oracle."""
    from unscene3d_b200.engine import coords as C

    outs = []
    for min_rows in (0, 1):
        C.set_row_ordering(min_rows)
        try:
            c = random_scene(4000, 31, batch=2, extent=28)
            x, y, fx, fy = _pair(eng, ora, c, cin)
            mx = eng.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, bias=False, dimension=3).cuda()
            my = ora.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, bias=False, dimension=3)
            torch.manual_seed(4)
            with torch.no_grad():
                mx.kernel.normal_(0, 0.05)
            _sync_params(mx, my)
            ox, oy = mx(x), my(y)
            g = torch.randn(oy.F.shape, generator=torch.Generator().manual_seed(9))
            ox.F.backward(g.cuda())
            oy.F.backward(g)
            assert rel_err(ox.F, oy.F) < 5e-5 and rel_err(fx.grad, fy.grad) < 5e-5 and rel_err(mx.kernel.grad, my.kernel.grad) < 5e-5
            outs.append((ox.F.detach().clone(), fx.grad.clone(), mx.kernel.grad.clone()))
        finally:
            C.set_row_ordering(32768)
    # (not bit-identical at this size: the unordered run of a small map deals its offsets to several CTAs that meet in
    # fp32 atomics; per row both runs add the same products)
    assert rel_err(outs[1][0], outs[0][0]) < 1e-5
    assert rel_err(outs[1][1], outs[0][1]) < 1e-5
    assert rel_err(outs[1][2], outs[0][2]) < 1e-5


@pytest.mark.parametrize("n,cin,cout,order_rows", [(60000, 32, 96, 0), (60000, 32, 96, 1 << 30), (4000, 64, 64, 1 << 30), (131, 16, 256, 1 << 30),
                                                  (40001, 96, 32, 0)])
def test_batchnorm_statistics_from_the_convolution_epilogue(eng, n, cin, cout, order_rows):
    """us3d_spconv_gather_mt_bn: mean / invstd / running statistics / counter produced by the convolution launch itself (epilogue
    of the range mode on large and pattern-ordered maps, the partial-tile reduction on small maps) equal those the statistics
    kernel computes from the written y (same sums, other order: 1e-5), including a ragged last tile; the scratch is left zero."""
    from unscene3d_b200.engine import coords as C
    from unscene3d_b200.engine import functional as Fn

    C.set_row_ordering(order_rows)
    try:
        c = random_scene(n, 77, batch=2, extent=64)
        x = eng.SparseTensor(torch.zeros(c.shape[0], 1, device="cuda"), torch.from_numpy(c).cuda())
        cm, key = x.coordinate_manager, x.coordinate_map_key
        table = cm.forward_table(key, key, (3, 3, 3))
        g = torch.Generator(device="cuda").manual_seed(5)
        f = torch.randn(table.n_rows, cin, device="cuda", generator=g) + 0.3
        w = torch.randn(27, cin, cout, device="cuda", generator=g) * 0.05
        rm0, rv0 = torch.randn(cout, device="cuda", generator=g), torch.rand(cout, device="cuda", generator=g) + 0.5
        res = []
        for fused in (True, False):
            Fn.set_fused_bn_stats(fused)
            rm, rv, nbt = rm0.clone(), rv0.clone(), torch.tensor(3, device="cuda")
            req = Fn.BnRequest(rm, rv, 0.1, 1e-5, nbt)
            y = Fn.spconv_gather(f, table, w, cin, cout, False, False, bn=req)
            res.append((y, req.mean.clone(), req.invstd.clone(), rm, rv, int(nbt)))
        # y: same products; on large maps the launch with statistics keeps two accumulator sets and the unfused weight operand,
        # i.e. another fp32 summation order of the three split terms
        assert rel_err(res[0][0], res[1][0]) < 1e-5
        ref_mean, ref_var = res[1][0].double().mean(0), res[1][0].double().var(0, unbiased=False)
        assert rel_err(res[1][1], ref_mean) < 1e-5 and rel_err(res[1][2], (ref_var + 1e-5).rsqrt()) < 1e-5
        for a, b in zip(res[0][1:5], res[1][1:5]):
            assert rel_err(a, b) < 1e-5
        assert res[0][5] == res[1][5] == 4
        assert float(Fn._bn_workspace(f.device, cout)[: 2 * cout + 1].abs().max()) == 0.0
    finally:
        Fn.set_fused_bn_stats(True)
        C.set_row_ordering(32768)


def test_launch_lists_equal_the_call_by_call_route(eng):
    """engine/blocks.py: a residual block's launches issued by one us3d_run_ops call are the launches of the call-by-call route
    (same kernels, arguments, order): features, BatchNorm buffers and gradients agree to the run-to-run noise of the atomics in the
    statistics / weight-gradient sums (1e-6 / 1e-5), in training and in evaluation mode, with and without a shortcut convolution."""
    from unscene3d_b200 import _lib, models
    from unscene3d_b200.engine import blocks as B
    from unscene3d_b200.utils import BackboneConfig, seeded_state

    c = random_scene(9000, 3, batch=2, extent=48)
    coords = torch.from_numpy(c).cuda()
    feats = torch.randn(c.shape[0], 3, generator=torch.Generator().manual_seed(1)).cuda()
    res = {}
    for on in (True, False, None):  # None: the call-by-call route a second time = the run-to-run noise floor of the atomics
        B.set_launch_lists(bool(on))
        try:
            # Res16UNet34C: stages of 2-6 blocks (one launch list per stage), weight images packed lazily DURING list assembly
            net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
            net.load_state_dict(seeded_state(net, 0))
            net = net.cuda().train()
            _lib.reset_launch_count()
            out, _ = net(eng.SparseTensor(feats, coords))
            (out.F * torch.linspace(-1, 1, out.F.shape[1], device="cuda")).mean().backward()
            grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
            bufs = {k: b.detach().clone() for k, b in net.named_buffers()}
            net.eval()
            with torch.no_grad():
                ev, _ = net(eng.SparseTensor(feats, coords))
            res[on] = (out.F.detach().clone(), grads, bufs, ev.F.clone(), _lib.launch_count())
        finally:
            B.set_launch_lists(True)
    # same kernels except the shortcut gradient: the list route lets the input-gradient launch accumulate it (y += ...), the
    # call-by-call route adds it in a separate pass — one launch per residual block with an input gradient
    n_blocks = sum(1 for m in net.modules() if type(m).__name__ == "BasicBlock")
    assert 0 <= res[False][4] - res[True][4] <= n_blocks, (res[False][4], res[True][4], n_blocks)
    # the statistics' shared-memory / fp64 atomics and the weight gradient's reds make two runs of ONE route differ in the last
    # bits (amplified to ~1e-5 by 30 layers); the two routes must differ by no more than two such runs do
    floor_f = max(rel_err(res[None][0], res[False][0]), rel_err(res[None][3], res[False][3]), 1e-7)
    assert floor_f < 2e-4
    assert rel_err(res[True][0], res[False][0]) < 4 * floor_f and rel_err(res[True][3], res[False][3]) < 4 * floor_f
    assert set(res[True][1]) == set(res[False][1])
    for k, g in res[False][1].items():
        floor_g = max(rel_err(res[None][1][k], g), 1e-6)
        assert rel_err(res[True][1][k], g) < 8 * floor_g + 1e-5, k
    for k, b in res[False][2].items():
        assert rel_err(res[True][2][k].float(), b.float()) < 1e-5, k


def test_transposed_convolution_on_pattern_ordered_tables(eng, ora):
    from unscene3d_b200.engine import coords as C

    C.set_row_ordering(1)
    try:
        test_transposed_convolution_forward_backward(eng, ora, 128, 96)
    finally:
        C.set_row_ordering(32768)


def test_pack_network_images_equal_the_per_layer_images(eng):
    """One launch packs every convolution of a network; byte-identical to the per-layer packing, including the column slices
    of the 384-channel decoder convolutions, and marked current so that the convolutions do not pack again."""
    from unscene3d_b200 import models
    from unscene3d_b200.engine import functional as Fn
    from unscene3d_b200.utils import BackboneConfig, seeded_state

    net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
    net.load_state_dict(seeded_state(net, 0))
    net = net.cuda()
    Fn.pack_network(net)
    seen = wide = 0
    for m in net.modules():
        if not (hasattr(m, "kernel_volume") and hasattr(m, "IS_TRANSPOSE")) or m.kernel.shape[-2] <= 4:
            continue
        if not hasattr(m.kernel, "_us3d_packs"):  # shapes the tensor-core kernels do not take (`final`: 96 -> 20)
            assert m.kernel.shape[-1] % 16 != 0
            continue
        key, fwd, bwd = m.kernel._us3d_packs
        fwd, bwd = fwd.clone(), [(c0, nc, img.clone()) for c0, nc, img in bwd]
        w3 = m.kernel.detach().view(m.kernel_volume, m.kernel.shape[-2], m.kernel.shape[-1])
        # the plan's images are current: no packing on use
        before = unscene3d_launches()
        f2, b2 = Fn.packed_weights(m.kernel, w3, key[3], 3, True)
        assert unscene3d_launches() == before and f2.data_ptr() == m.kernel._us3d_packs[1].data_ptr()
        del m.kernel._us3d_packs
        f3, b3 = Fn.packed_weights(m.kernel, w3, key[3], 3, True)   # per-layer packing
        assert torch.equal(fwd, f3)
        assert len(bwd) == len(b3)
        for (c0, nc, img), (d0, dn, img3) in zip(bwd, b3):
            assert (c0, nc) == (d0, dn) and torch.equal(img, img3)
        seen += 1
        wide += len(bwd) > 1
    assert seen >= 60 and wide >= 1


def unscene3d_launches():
    from unscene3d_b200 import _lib

    return _lib.launch_count()


def test_empty_and_single_voxel(eng, ora):
    c = np.array([[0, 5, -3, 2]], dtype=np.int32)
    x, y, fx, fy = _pair(eng, ora, c, 8)
    mx = eng.MinkowskiConvolution(8, 8, kernel_size=3, dimension=3).cuda()
    my = ora.MinkowskiConvolution(8, 8, kernel_size=3, dimension=3)
    _sync_params(mx, my)
    assert rel_err(mx(x).F, my(y).F) < 1e-6
    px = eng.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(x)
    assert px.F.shape == (1, 8) and rel_err(px.F, fy) < 1e-6


# --------------------------------------------------------------------------------------------- BN & friends
@pytest.mark.parametrize("n,c,relu,res", [(5000, 32, False, False), (3001, 96, True, False), (777, 256, True, True), (64, 3, False, True)])
def test_batchnorm_train_forward_backward(eng, n, c, relu, res):
    from unscene3d_b200.engine import functional as Fn

    torch.manual_seed(0)
    x = (torch.randn(n, c) * 2 + 0.5)
    r = torch.randn(n, c) if res else None
    bn = torch.nn.BatchNorm1d(c, momentum=0.02)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.uniform_(-0.3, 0.3)
    xr = x.clone().requires_grad_()
    rr = r.clone().requires_grad_() if res else None
    ref = bn(xr)
    if res:
        ref = ref + rr
    if relu:
        ref = torch.relu(ref)
    g = torch.randn(n, c)
    ref.backward(g)

    xg = x.clone().cuda().requires_grad_()
    rg = r.clone().cuda().requires_grad_() if res else None
    w = bn.weight.detach().clone().cuda().requires_grad_()
    b = bn.bias.detach().clone().cuda().requires_grad_()
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    out = Fn.BatchNormFunction.apply(xg, w, b, rg, rm, rv, 0.02, 1e-5, True, relu)
    out.backward(g.cuda())
    assert rel_err(out, ref) < 1e-5
    assert rel_err(xg.grad, xr.grad) < 1e-4
    assert rel_err(w.grad, bn.weight.grad) < 1e-4
    assert rel_err(b.grad, bn.bias.grad) < 1e-4
    if res:
        assert rel_err(rg.grad, rr.grad) < 1e-6
    assert rel_err(rm, bn.running_mean) < 1e-5 and rel_err(rv, bn.running_var) < 1e-5


def test_fused_batchnorm_planes_counters_and_workspace(eng):
    """The fused BatchNorm passes: bf16 planes written by the apply / backward pass are bit-identical to the stand-alone
    split kernel's and are what the neighbouring convolutions consume (no split launch between conv-bn-relu-conv);
    num_batches_tracked is incremented on the device; the shared workspace is zero again after every pass."""
    from unscene3d_b200 import _lib
    from unscene3d_b200.engine import functional as Fn

    torch.manual_seed(3)
    c = random_scene(4000, 5, batch=2, extent=28)
    feats = torch.randn(c.shape[0], 32)
    net = torch.nn.Sequential(eng.MinkowskiConvolution(32, 64, kernel_size=3, dimension=3), eng.MinkowskiBatchNorm(64),
                              eng.MinkowskiReLU(), eng.MinkowskiConvolution(64, 96, kernel_size=3, dimension=3),
                              eng.MinkowskiBatchNorm(96), eng.MinkowskiReLU()).cuda().train()
    x = eng.SparseTensor(feats.cuda().requires_grad_(), torch.from_numpy(c).cuda())
    calls = {"n": 0}
    real = _lib.lib.us3d_split_bf16

    class Spy:
        def __getattr__(self, name):
            fn = getattr(_lib.lib, name)
            if name != "us3d_split_bf16":
                return fn

            def counted(*a):
                calls["n"] += 1
                return real(*a)
            return counted

    Fn.lib = Spy()
    try:
        h = net[2](net[1](net[0](x)))
        mid = h.F
        planes = getattr(mid, "_us3d_planes", None)
        assert planes is not None and planes[2] == mid._version, "BatchNorm apply did not leave bf16 planes on its output"
        hi = torch.empty_like(planes[0]); lo = torch.empty_like(planes[1])
        _lib.check(real(mid.data_ptr(), mid.shape[1], mid.shape[0], mid.shape[1], hi.data_ptr(), lo.data_ptr(), Fn._stream()))
        assert torch.equal(hi.view(torch.int16), planes[0].view(torch.int16)) and torch.equal(lo.view(torch.int16), planes[1].view(torch.int16))
        before = calls["n"]
        out = net[5](net[4](net[3](h))).F
        assert calls["n"] == before, "the second convolution re-split an activation that already had planes"
        out.backward(torch.randn_like(out))
        # backward: dy of both convolutions comes from a BatchNorm backward pass that wrote its planes
        assert calls["n"] == before, f"{calls['n'] - before} split launches in backward"
    finally:
        Fn.lib = _lib.lib
    assert int(net[1].bn.num_batches_tracked) == 1 and int(net[4].bn.num_batches_tracked) == 1
    ws = Fn._bn_workspace(mid.device, 96)
    assert float(ws.abs().max()) == 0.0, "BatchNorm workspace not returned to zero"


def test_coordinate_stream_gives_identical_maps_and_features(eng):
    """Coordinate maps built on a dedicated stream (engine.set_coordinate_stream) while the compute stream is busy:
    same maps, same kernel maps, same convolution output as the single-stream path, step after step without a
    device-wide synchronisation in between."""
    torch.manual_seed(11)
    conv = torch.nn.Sequential(eng.MinkowskiConvolution(16, 32, kernel_size=3, dimension=3), eng.MinkowskiBatchNorm(32), eng.MinkowskiReLU(),
                               eng.MinkowskiConvolution(32, 32, kernel_size=2, stride=2, dimension=3)).cuda().train()
    scenes = [random_scene(6000 + 500 * i, 40 + i, batch=2, extent=30) for i in range(3)]
    feats = [torch.randn(c.shape[0], 16).cuda() for c in scenes]
    coords = [torch.from_numpy(c).cuda() for c in scenes]

    def run():
        outs = []
        busy = torch.randn(4096, 4096, device="cuda")
        for c, f in zip(coords, feats):
            for _ in range(3):
                busy = busy @ busy.clamp(-1e-3, 1e-3)  # keeps the compute stream occupied while the next maps are built
            y = conv(eng.SparseTensor(f, c))
            outs.append((y.C.clone(), y.F.detach().clone()))
        torch.cuda.synchronize()
        return outs

    ref = run()
    eng.set_coordinate_stream(torch.cuda.Stream(priority=-1))
    try:
        got = run()
    finally:
        eng.set_coordinate_stream(None)
    for (c0, f0), (c1, f1) in zip(ref, got):
        assert torch.equal(c0, c1)
        assert rel_err(f1, f0.cpu()) < 1e-6


def test_batchnorm_eval_mode(eng):
    torch.manual_seed(1)
    m = eng.MinkowskiBatchNorm(16, momentum=0.1).cuda()
    m.bn.running_mean.uniform_(-1, 1)
    m.bn.running_var.uniform_(0.5, 2)
    m.eval()
    c = random_scene(500, 3)
    f = torch.randn(c.shape[0], 16)
    x = eng.SparseTensor(f.cuda().requires_grad_(), torch.from_numpy(c).cuda())
    ref_bn = torch.nn.BatchNorm1d(16).eval()
    ref_bn.load_state_dict({k: v.cpu() for k, v in m.bn.state_dict().items()})
    fr = f.clone().requires_grad_()
    ref = ref_bn(fr)
    out = m(x).F
    g = torch.randn_like(ref)
    out.backward(g.cuda())
    ref.backward(g)
    assert rel_err(out, ref) < 1e-5
    assert rel_err(x.F.grad, fr.grad) < 1e-5


def test_relu_add_cat_pool(eng, ora):
    c = random_scene(4000, 31, batch=2, extent=20)
    x, y, fx, fy = _pair(eng, ora, c, 12)
    x2, y2 = x._like(fx * 0.5 + 1), y._like(fy * 0.5 + 1)
    ox = eng.MinkowskiReLU(inplace=True)(eng.cat(x, x2) + eng.cat(x2, x))
    oy = ora.MinkowskiReLU()(ora.cat(y, y2) + ora.cat(y2, y))
    for mode in ("Avg", "Sum", "Max"):
        px = getattr(eng, f"Minkowski{mode}Pooling")(kernel_size=2, stride=2, dimension=3)(ox)
        py = getattr(ora, f"Minkowski{mode}Pooling")(kernel_size=2, stride=2, dimension=3)(oy)
        assert torch.equal(px.C.cpu(), py.C)
        assert rel_err(px.F, py.F) < 1e-6
    px = eng.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(ox)
    py = ora.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)(oy)
    g = torch.randn_like(py.F)
    px.F.backward(g.cuda())
    py.F.backward(g)
    assert rel_err(fx.grad, fy.grad) < 1e-6
    # decomposition by batch index keeps row order
    for a, b in zip(px.decomposed_features, py.decomposed_features):
        assert rel_err(a, b) < 1e-6
    for a, b in zip(px.decomposed_coordinates, py.decomposed_coordinates):
        assert torch.equal(a.cpu(), b)


# --------------------------------------------------------------------------------------------- decoder helpers
# n chosen to cover every residency tier of the cluster kernel: one CTA (<= 8192 rows), 2..16 CTAs with all rows in
# registers (<= 131072), rows in shared memory (<= 327680), rows streamed from L2 beyond that
@pytest.mark.parametrize("n,m,span", [(1, 1, 8), (3, 7, 8), (5, 3, 8), (31, 10, 8), (100, 100, 8), (777, 100, 8), (5000, 100, 8), (20000, 60, 20),
                                      (131072, 40, 40), (200000, 100, 150), (340000, 30, 60)])
def test_fps_indices_bit_exact_on_integer_coords(eng, n, m, span):
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn

    rng = np.random.default_rng(n)
    pts = rng.integers(-span, span + 1, size=(2, n, 3)).astype(np.float32)  # many exact distance ties, some |p|^2 = 0
    temp_probe = n >= 20000
    got = Fn.furthest_point_sampling(torch.from_numpy(pts).cuda(), m).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], ops_cpu.furthest_point_sampling(pts[b], m))
    if temp_probe:  # run-to-run identical (no atomics, fixed reduction order)
        again = Fn.furthest_point_sampling(torch.from_numpy(pts).cuda(), m).cpu().numpy()
        assert np.array_equal(got, again)


def test_fps_all_rows_skipped_returns_row_zero(eng):
    """Rows with |p|^2 <= 1e-3 never win (sampling_gpu.cu:103-104): a scene of such rows yields index 0 throughout."""
    from unscene3d_b200.engine import functional as Fn

    pts = torch.zeros(1, 700, 3).cuda()
    assert torch.equal(Fn.furthest_point_sampling(pts, 7).cpu(), torch.zeros(1, 7, dtype=torch.int32))


@pytest.mark.parametrize("S,Q,T_all,T,tgt_dtype,weighted", [(1500, 100, 20, 20, torch.bool, False), (333, 100, 30, 7, torch.float32, True),
                                                             (64, 10, 5, 0, torch.bool, False), (2000, 100, 25, 25, torch.uint8, True)])
def test_mask_losses_match_oracle(eng, S, Q, T_all, T, tgt_dtype, weighted):
    """Set-criterion mask losses of one scene (models/criterion.py:22-73, 176-216): forward and the gradient with respect to
    the mask logits against the oracle's restatement evaluated in float64; 1e-5 relative."""
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn

    g = torch.Generator().manual_seed(S + T)
    logits = torch.randn(S, Q, generator=g) * 3
    tgt = (torch.rand(T_all, S, generator=g) < 0.2).to(tgt_dtype)
    qidx = torch.randperm(Q, generator=g)[:T]
    tidx = torch.randperm(T_all, generator=g)[:T]
    w = (torch.rand(T, generator=g) < 0.7).float() if weighted else None
    n = max(T, 1)
    ref_in = logits.clone().double().requires_grad_()
    r_ce, r_dice = ops_cpu.mask_losses(ref_in, tgt.double(), qidx, tidx, None if w is None else w.double(), n)
    gu = torch.tensor([0.7, -1.3])
    (gu[0] * r_ce + gu[1] * r_dice).backward()
    cu_in = logits.clone().cuda().requires_grad_()
    c_ce, c_dice = Fn.mask_losses(cu_in, tgt.cuda(), qidx.cuda(), tidx.cuda(), None if w is None else w.cuda(), n)
    (gu[0] * c_ce + gu[1] * c_dice).backward()
    assert abs(float(c_ce) - float(r_ce)) <= 1e-5 * max(abs(float(r_ce)), 1e-6)
    assert abs(float(c_dice) - float(r_dice)) <= 1e-5 * max(abs(float(r_dice)), 1e-6)
    if T:
        assert rel_err(cu_in.grad, ref_in.grad) < 1e-5
    else:
        assert float(cu_in.grad.abs().max()) == 0.0


@pytest.mark.parametrize("steps", [1, 2, 4])
def test_segment_attention_masks_equal_gather_pool_threshold(eng, steps):
    """Decoder attention masks (models/mask3d.py:419-446): the sparse product sigmoid(A seg) < 0.5 against the reference's
    sequence run with this engine's own kernels (gather to the voxels, concatenate, MinkowskiAvgPooling x steps, threshold).
    Both are fp32 sums in different orders: entries whose pooled logit lies within 1e-5 of zero may differ, nothing else."""
    from unscene3d_b200.engine import functional as Fn

    c = random_scene(6000, 41, batch=3, extent=40)
    g = torch.Generator().manual_seed(1)
    x = eng.SparseTensor(torch.zeros(c.shape[0], 1, device="cuda"), torch.from_numpy(c).cuda())
    Q = 100
    p2s, seg = [], []
    for coords_b in x.decomposed_coordinates:
        S = 37 + len(p2s) * 11
        ids = torch.randint(0, S, (coords_b.shape[0],), generator=g)
        ids[:S] = torch.arange(S)
        p2s.append(ids.cuda())
        seg.append((torch.randn(S, Q, generator=g) * 2).cuda())
    key, bits = Fn.segment_attention_masks(x, seg, p2s, steps)
    t = eng.SparseTensor(torch.cat([s[p] for s, p in zip(seg, p2s)]), coordinate_map_key=x.coordinate_map_key,
                         coordinate_manager=x.coordinate_manager)
    pool = eng.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)
    for _ in range(steps):
        t = pool(t)
    assert t.coordinate_map_key == key and bits.shape == t.F.shape and bits.dtype == torch.bool
    want = t.F.sigmoid() < 0.5
    differ = want != bits
    assert int(differ.sum()) <= 1e-4 * bits.numel()
    assert float(t.F[differ].abs().max()) < 1e-5 if bool(differ.any()) else True
    # second call re-uses the cached matrix
    key2, bits2 = Fn.segment_attention_masks(x, seg, p2s, steps)
    assert key2 == key and torch.equal(bits2, bits)


def test_segment_mean(eng):
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn

    torch.manual_seed(0)
    src = torch.randn(20000, 128)
    idx = torch.randint(0, 900, (20000,))
    idx[0] = 899
    ref_in = src.clone().requires_grad_()
    ref = ops_cpu.scatter_mean(ref_in, idx)
    s = src.clone().cuda().requires_grad_()
    out = Fn.SegmentMeanFunction.apply(s, idx.cuda(), 900)
    g = torch.randn_like(ref)
    out.backward(g.cuda())
    ref.backward(g)
    assert rel_err(out, ref) < 1e-5
    assert rel_err(s.grad, ref_in.grad) < 1e-6


def test_matcher_cost(eng):
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn

    torch.manual_seed(0)
    S, Q, T = 1500, 100, 20
    logits = torch.randn(Q, 3)
    masks = torch.randn(S, Q) * 4
    tgt = torch.rand(T, S) < 0.1
    labels = torch.ones(T, dtype=torch.long)
    labels[3] = 253
    ref = ops_cpu.matcher_cost(logits, masks, tgt, labels, 2.0, 5.0, 2.0)
    got = Fn.matcher_cost(masks.cuda(), tgt.float().cuda(), logits.softmax(-1).cuda(), labels.cuda(), 2.0, 5.0, 2.0)
    assert rel_err(got, ref) < 1e-5
    from scipy.optimize import linear_sum_assignment

    assert [a.tolist() for a in linear_sum_assignment(got.cpu())] == [a.tolist() for a in linear_sum_assignment(ref)]


def test_matcher_cost_rejects_labels_outside_the_class_range_and_takes_cpu_labels(eng):
    """The reference indexes `out_prob[:, tgt_ids]` (models/matcher.py:122-127): CPU or CUDA labels alike, IndexError for a label
    outside [0, num_classes).  Here: CPU labels are moved, an invalid label poisons its cost column with NaN (the assignment
    solver then rejects the matrix) and nothing is read out of bounds."""
    from unscene3d_b200.engine import functional as Fn

    torch.manual_seed(1)
    S, Q, T = 300, 100, 6
    prob = torch.randn(Q, 3).softmax(-1).cuda()
    masks = (torch.randn(S, Q) * 4).cuda()
    tgt = (torch.rand(T, S) < 0.2).float().cuda()
    labels = torch.tensor([1, 0, 2, 253, 1, 2])
    good = Fn.matcher_cost(masks, tgt, prob, labels, 2.0, 5.0, 2.0)          # CPU labels
    same = Fn.matcher_cost(masks, tgt, prob, labels.cuda(), 2.0, 5.0, 2.0)
    assert torch.equal(good, same) and bool(torch.isfinite(good).all())
    labels[4] = 255                                                           # the ignore value / a mis-set num_classes
    bad = Fn.matcher_cost(masks, tgt, prob, labels, 2.0, 5.0, 2.0)
    assert bool(torch.isnan(bad[:, 4]).all()) and bool(torch.isfinite(bad[:, [0, 1, 2, 3, 5]]).all())
    from scipy.optimize import linear_sum_assignment

    with pytest.raises(ValueError):
        linear_sum_assignment(bad.cpu())


@pytest.mark.parametrize("n,d_out,normalize", [(1, 64, True), (777, 64, True), (20000, 64, True), (500, 16, False)])
def test_fourier_posenc_matches_oracle(eng, n, d_out, normalize):
    """us3d_fourier_posenc against the oracle's line-by-line restatement of models/position_embedding.py:12-40, 128-160 (which
    the Mask3D golden pins to the unmodified reference file).  fp32: the projection is a 3-term dot product of values up to
    2 pi, sin / cos of arguments up to ~20 -> absolute agreement 2e-6."""
    from oracle import ops_cpu
    from unscene3d_b200.engine import functional as Fn
    from unscene3d_b200.models.position_embedding import PositionEmbeddingCoordsSine

    torch.manual_seed(n)
    xyz = (torch.rand(n, 3) * torch.tensor([5.0, 4.0, 2.4]) - torch.tensor([2.5, 2.0, 0.1]))
    xyz = torch.floor(xyz / 0.02) * 0.02
    mod = PositionEmbeddingCoordsSine(pos_type="fourier", d_pos=128, gauss_scale=1.0, normalize=normalize)
    lo, hi = (xyz.min(0)[0], xyz.max(0)[0] + (1.0 if n == 1 else 0.0)) if normalize else (None, None)
    want = ops_cpu.fourier_posenc(xyz, mod.gauss_B, d_out, lo, hi)
    got = Fn.fourier_posenc(xyz.cuda(), mod.gauss_B.cuda(), d_out, None if lo is None else lo.cuda(), None if hi is None else hi.cuda())
    assert got.shape == (n, 2 * d_out)
    assert float((got.cpu() - want).abs().max()) < 2e-6
    # module surface: [B, d_pos, N] like the reference, rows through encode_rows
    m = mod.cuda()
    full = m(xyz.cuda()[None], num_channels=2 * d_out, input_range=None if lo is None else [lo.cuda()[None], hi.cuda()[None]])
    assert full.shape == (1, 2 * d_out, n) and torch.equal(full[0].permute(1, 0), got)


def test_torch_scatter_shim_cpu_tensors_and_dim_size_validation(eng):
    """torch_scatter.scatter_mean as the reference calls it: CUDA rows on the hot path (models/mask3d.py:223), CPU tensors in
    the evaluation post-processing (trainer/trainer.py:449); an index tensor on another device is moved; an explicit dim_size
    that the indices exceed raises like torch_scatter does."""
    import torch_scatter  # the shim (unscene3d_b200/shims is on sys.path once the package is imported)
    from oracle import ops_cpu

    assert "unscene3d_b200" in torch_scatter.__file__
    torch.manual_seed(2)
    src = torch.randn(5000, 7)
    idx = torch.randint(0, 40, (5000,))
    idx[0] = 39
    want = ops_cpu.scatter_mean(src, idx)
    assert rel_err(torch_scatter.scatter_mean(src, idx, dim=0), want) < 1e-6          # CPU in, CPU out
    got = torch_scatter.scatter_mean(src.cuda(), idx, dim=0)                            # CUDA rows, CPU index
    assert got.is_cuda and rel_err(got, want) < 1e-5
    big = torch_scatter.scatter_mean(src.cuda(), idx.cuda(), dim=0, dim_size=64)
    assert big.shape == (64, 7) and float(big[40:].abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        torch_scatter.scatter_mean(src.cuda(), idx.cuda(), dim=0, dim_size=10)


def test_cpu_tensor_is_rejected_loudly(eng):
    with pytest.raises(RuntimeError):
        eng.SparseTensor(torch.zeros(3, 1), torch.zeros(3, 4, dtype=torch.int32))


def test_deferred_batchnorm_fusion_and_no_grad_access(eng, ora):
    """conv -> bn -> (+= residual) -> relu is ONE fused apply pass; touching metadata or the features inside
    torch.no_grad() (as models/mask3d.py:205-209 does with aux[-1]) must not cut the autograd graph."""
    from unscene3d_b200 import _lib

    c = random_scene(2000, 41, batch=2, extent=20)
    x, y, fx, fy = _pair(eng, ora, c, 16)
    bx, by = eng.MinkowskiBatchNorm(16, momentum=0.1).cuda(), ora.MinkowskiBatchNorm(16, momentum=0.1)
    rx, ry = eng.MinkowskiReLU(inplace=True), ora.MinkowskiReLU(inplace=True)
    ox = bx(x)
    before = _lib.launch_count()
    ox += x
    ox = rx(ox)
    assert _lib.launch_count() == before, "residual add and ReLU must stay deferred"
    with torch.no_grad():
        assert ox.device.type == "cuda" and ox.shape == (c.shape[0], 16)
        _ = ox.F  # materialises here, under no_grad
    assert ox.F.requires_grad and ox.F.grad_fn is not None
    oy = by(y)
    oy += y
    oy = ry(oy)
    assert rel_err(ox.F, oy.F) < 1e-5
    g = torch.randn_like(oy.F)
    ox.F.backward(g.cuda())
    oy.F.backward(g)
    assert rel_err(fx.grad, fy.grad) < 1e-4
    assert rel_err(bx.bn.weight.grad, by.bn.weight.grad) < 1e-4
