"""Backbone parity.

CPU (`-m "not gpu"`): our model definitions over the CPU oracle reproduce the golden vectors that
the UNMODIFIED reference model files produced over the same oracle (tests/golden/make_golden.py) —
this pins unscene3d_b200/models/*.py to /root/reference/models/*.py; when the reference tree is
present the two are also compared live (state-dict names, shapes, BN momenta, outputs).

GPU (`-m gpu`): the same model on the CUDA backend (libus3d through the C ABI) against the golden
vectors and against the oracle run live — forward features, loss, parameter gradients, BN buffers.
Tolerance: 1e-3 relative (north_star) for the model-level comparisons; the fp32 SIMT kernels land
around 1e-6.
"""
import os

import numpy as np
import pytest
import torch

from golden.make_golden import CASES, GRAD_KEYS, case_inputs, run_case
from helpers import (Cfg, deterministic_state, have_reference, our_models_on_oracle, random_scene,
                     reference_models_on_oracle)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))


def compare(res, gold, rtol, grad_rtol=None):
    """Forward quantities (feature samples, sums, loss, BN running statistics) are held to `rtol`.
    Parameter gradients get `grad_rtol`: a gradient is a sum over ~10^6 elements gated by ReLU masks, and an
    implementation whose forward differs from the oracle by 1e-6..4e-5 (different fp32 summation order; the
    tensor-core path adds ~1e-5 per layer from the bf16 split and TMEM accumulation) flips a handful of
    those masks; each flip is a 100 % error on one element, so the L2 error of a gradient is
    ~sqrt(#flips / #elements) — measured 3e-3 with the exact-fp32 kernels and 1.4e-2 with the tensor-core
    kernels — no matter how exact the backward kernels are (the kernels themselves agree with the oracle to
    5e-5 on identical inputs, tests/test_gpu_ops.py).  Hence 5e-2 here plus a cosine-similarity bound."""
    grad_rtol = grad_rtol or rtol
    for k, g in gold.items():
        r = res[k]
        if k.startswith(("shape", "idx")):
            assert np.array_equal(r, g), k
        else:
            tol = grad_rtol if k.startswith(("grad:", "gnorm:")) else rtol
            scale = max(float(np.abs(g).max()), 1e-12)
            err = float(np.abs(np.asarray(r, dtype=np.float64) - g).max()) / scale
            assert err < tol, f"{k}: max error {err:.3e} (relative to max |golden|) exceeds {tol}"


@pytest.mark.parametrize("name", list(CASES))
def test_our_models_on_oracle_match_reference_golden(name):
    from oracle import me_cpu

    res = run_case(our_models_on_oracle(), me_cpu, name)
    compare(res, load_golden(name), rtol=1e-6)


@pytest.mark.skipif(not have_reference(), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("cls", ["Res16UNet14", "Res16UNet18", "Res16UNet34", "Res16UNet34C", "Res16UNet14A", "Res16UNet34CMultiRes"])
def test_state_dict_and_outputs_identical_to_reference(cls):
    from oracle import me_cpu

    ref = reference_models_on_oracle()
    import models.res16unet as R

    ours = our_models_on_oracle().res16unet
    a = getattr(R, cls)(3, 20, Cfg(), D=3, out_fpn=True)
    b = getattr(ours, cls)(3, 20, Cfg(), D=3, out_fpn=True)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape for k in sa)
    mom = lambda m: {n: x.momentum for n, x in m.named_modules() if isinstance(x, torch.nn.BatchNorm1d)}
    assert mom(a) == mom(b)
    st = deterministic_state(a, 5)
    a.load_state_dict(st)
    b.load_state_dict(st)
    c = random_scene(1500, 9, batch=2)
    f = torch.randn(c.shape[0], 3)
    oa = a(me_cpu.SparseTensor(f, torch.from_numpy(c)))
    ob = b(me_cpu.SparseTensor(f, torch.from_numpy(c)))
    assert torch.equal(oa[0].F, ob[0].F)
    fa = list(oa[1].values()) if isinstance(oa[1], dict) else oa[1]
    fb = list(ob[1].values()) if isinstance(ob[1], dict) else ob[1]
    for u, v in zip(fa, fb):
        assert torch.equal(u.F, v.F) and torch.equal(u.C, v.C)


def _backbone_step(net, me, f, c, w, dtype=torch.float32):
    out, aux = net(me.SparseTensor(f.to(dtype), c))
    (out.F * w.to(out.F)).mean().backward()
    grads = {k: p.grad.detach().double().cpu().clone() for k, p in net.named_parameters() if p.grad is not None}
    net.zero_grad(set_to_none=True)
    return out.F.detach().double().cpu(), grads


def test_relu_mask_flips_explain_the_gradient_tolerance_fp32_vs_fp64_oracle():
    """Pins the argument behind the gradient tolerances: the SAME oracle run in fp32 and in fp64 (features agree to 1e-6)
    disagrees on parameter gradients by far more than that, because a few near-zero pre-activations change sign; with the
    fp32 run's ReLU masks replayed in the fp64 run the gradients agree at rounding level.  The CUDA tests use the same replay
    to hold the gradients of the tensor-core path to 1e-3 (test_cuda_backbone_gradients_with_relu_masks_replayed)."""
    from helpers import record_relu_masks, replay_relu_masks
    from oracle import me_cpu

    c = torch.from_numpy(random_scene(3000, 21, batch=2, extent=28))
    torch.manual_seed(4)
    f = torch.randn(c.shape[0], 3)
    w = torch.linspace(-1, 1, 96)
    R = our_models_on_oracle().res16unet
    net32 = R.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    st = deterministic_state(net32, 13)
    net32.load_state_dict(st)
    net64 = R.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True).double()
    net64.load_state_dict(st)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp(min=1e-300))
    with record_relu_masks(me_cpu) as masks:
        o32, g32 = _backbone_step(net32, me_cpu, f, c, w)
    net64.load_state_dict(st)
    o64, g64 = _backbone_step(net64, me_cpu, f, c, w, torch.float64)
    free = max(rel(g32[k], g64[k]) for k in g64)
    flips = []
    net64.load_state_dict(st)
    with replay_relu_masks(me_cpu, masks, flips):
        o64r, g64r = _backbone_step(net64, me_cpu, f, c, w, torch.float64)
    replayed = max(rel(g32[k], g64r[k]) for k in g64r)
    assert rel(o32, o64) < 1e-5
    assert replayed < 1e-4, replayed
    n_flip = sum(a for a, _ in flips)
    if n_flip:  # the usual case: a handful of flipped decisions cost orders of magnitude in gradient agreement
        assert free > 5 * replayed, (free, replayed, n_flip)


# ------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("mode", [3, 0])
def test_cuda_backbone_gradients_with_relu_masks_replayed(mode):
    """Whole-network backward at the north-star tolerance: the CUDA run (module-by-module route, which exposes every ReLU)
    records its ReLU masks, the oracle replays them; features, loss and EVERY parameter gradient agree to 1e-3 in the
    production arithmetic (mode 3, three-term bf16 split on tcgen05) and to 1e-4 with the exact-fp32 kernels (mode 0)."""
    import unscene3d_b200  # noqa: F401
    from helpers import record_relu_masks, replay_relu_masks
    from oracle import me_cpu
    from unscene3d_b200 import engine, models
    from unscene3d_b200.engine import blocks
    from unscene3d_b200.engine import functional as Fn

    c = torch.from_numpy(random_scene(6000, 78, batch=2, extent=36))
    torch.manual_seed(5)
    f = torch.randn(c.shape[0], 3)
    w = torch.linspace(-1, 1, 96)
    cpu_net = our_models_on_oracle().res16unet.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    st = deterministic_state(cpu_net, 12)
    cpu_net.load_state_dict(st)
    gpu_net = models.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    gpu_net.load_state_dict(st)
    gpu_net.cuda()
    default_on = blocks._enabled["on"]
    blocks.set_fused_blocks(False)
    Fn.set_precision(mode)
    try:
        with record_relu_masks(engine) as masks:
            og, gg = _backbone_step(gpu_net, engine, f.cuda(), c.cuda(), w)
    finally:
        blocks.set_fused_blocks(default_on)
        Fn.set_precision(3)
    flips = []
    with replay_relu_masks(me_cpu, masks, flips):
        oc, gc = _backbone_step(cpu_net, me_cpu, f, c, w)
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp(min=1e-300))
    tol = 1e-3 if mode == 3 else 1e-4
    assert rel(og, oc) < tol
    assert gg.keys() == gc.keys()
    worst_k, worst = max(((k, rel(gg[k], gc[k])) for k in gc), key=lambda kv: kv[1])
    assert worst < tol, f"{worst_k}: {worst:.2e} (mode {mode}, {sum(a for a, _ in flips)} flipped ReLU decisions replayed)"


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_backbone_matches_golden(name):
    import unscene3d_b200
    from unscene3d_b200 import engine, models

    res = run_case(models, engine, name, device="cuda")
    compare(res, load_golden(name), rtol=1e-3, grad_rtol=5e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("row_order_min", [32768, 1])
def test_cuda_backbone_matches_oracle_live_forward_backward(row_order_min):
    """row_order_min = 1 forces the neighbour-pattern row order that production applies to maps of >= 32768 rows."""
    import unscene3d_b200
    from unscene3d_b200 import engine

    engine.set_row_ordering(row_order_min)
    try:
        _backbone_live()
    finally:
        engine.set_row_ordering(32768)


def _backbone_live():
    import unscene3d_b200
    from oracle import me_cpu
    from unscene3d_b200 import engine, models

    c = random_scene(6000, 77, batch=2, extent=36)
    torch.manual_seed(3)
    f = torch.randn(c.shape[0], 3)
    cpu_net = our_models_on_oracle().res16unet.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    st = deterministic_state(cpu_net, 11)
    cpu_net.load_state_dict(st)
    gpu_net = models.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True)
    gpu_net.load_state_dict(st)
    gpu_net.cuda()
    oc, fc = cpu_net(me_cpu.SparseTensor(f, torch.from_numpy(c)))
    og, fg = gpu_net(engine.SparseTensor(f.cuda(), torch.from_numpy(c).cuda()))

    def rel(a, b):
        return float((a.double().cpu() - b.double()).norm() / b.double().norm())

    # stride-1 rows keep input order; coarse maps share first-occurrence order with the oracle
    for u, v in zip(fg, fc):
        assert torch.equal(u.C.cpu(), v.C)
        assert rel(u.F, v.F) < 1e-3
    w = torch.linspace(-1, 1, oc.F.shape[1])
    (oc.F * w).mean().backward()
    (og.F * w.cuda()).mean().backward()
    pc, pg = dict(cpu_net.named_parameters()), dict(gpu_net.named_parameters())
    errs = sorted(rel(pg[k].grad, pc[k].grad) for k in pc if k != "final.kernel" and k != "final.bias")
    # see compare(): ReLU-mask flips bound the L2 agreement of gradients at ~sqrt(#flips/#elements)
    assert errs[-1] < 5e-2, f"worst parameter-gradient relative error {errs[-1]:.3e}"
    cos = min(float(torch.nn.functional.cosine_similarity(pg[k].grad.double().cpu().flatten(), pc[k].grad.double().flatten(), dim=0))
              for k in pc if pc[k].grad is not None)
    assert cos > 0.998, f"worst gradient cosine similarity {cos}"
    bc, bg = dict(cpu_net.named_buffers()), dict(gpu_net.named_buffers())
    for k in bc:
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel(bg[k], bc[k]) < 1e-4, k
        if k.endswith("num_batches_tracked"):
            assert int(bg[k]) == int(bc[k])


@pytest.mark.gpu
@pytest.mark.parametrize("inplanes,planes,train", [(96, 96, True), (128, 96, True), (32, 64, False), (384, 256, True)])
def test_fused_residual_block_equals_the_module_route(inplanes, planes, train):
    """engine.blocks runs a BasicBlock (and a conv-norm-relu transition) as ONE autograd node issuing the same kernels in the
    same order as the module-by-module route: outputs, input gradient, every parameter gradient and the BatchNorm buffers
    agree to 1e-5 (not bit for bit: small maps combine partial sums with fp32 atomics, as do the weight gradients)."""
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200 import engine
    from unscene3d_b200.engine import blocks
    from unscene3d_b200.models.modules.common import conv, get_norm, NormType
    from unscene3d_b200.models.modules.resnet_block import BasicBlock

    import torch.nn as nn

    c = random_scene(5000, 80, batch=2, extent=30)
    torch.manual_seed(6)
    f = torch.randn(c.shape[0], inplanes).cuda()
    ds = None
    if inplanes != planes:
        ds = nn.Sequential(conv(inplanes, planes, kernel_size=1, stride=1, D=3), get_norm(NormType.BATCH_NORM, planes, 3, bn_momentum=0.1))
    block = BasicBlock(inplanes, planes, downsample=ds, bn_momentum=0.1, D=3).cuda()
    trans = conv(planes, planes, kernel_size=2, stride=2, D=3).cuda()
    tnorm = get_norm(NormType.BATCH_NORM, planes, 3, bn_momentum=0.02).cuda()
    relu = engine.MinkowskiReLU(inplace=True)
    mods = nn.ModuleList([block, trans, tnorm]).train(train)
    state = {k: v.clone() for k, v in mods.state_dict().items()}
    g = None
    runs = []
    default_on = blocks._enabled["on"]
    for fused in (True, False):
        blocks.set_fused_blocks(fused)
        try:
            mods.load_state_dict(state)
            x = f.clone().requires_grad_()
            h = block(engine.SparseTensor(x, torch.from_numpy(c).cuda()))
            t = blocks.fused_conv_norm_relu(trans, tnorm, h) if fused else None
            if t is None:
                assert not fused
                t = relu(tnorm(trans(h)))
            if g is None:
                g = torch.randn_like(t.F)
            (t.F * g).sum().backward()
            runs.append((t.F.detach().clone(), x.grad.clone(), {k: p.grad.clone() for k, p in mods.named_parameters()},
                         {k: b.clone() for k, b in mods.named_buffers()}))
            mods.zero_grad(set_to_none=True)
        finally:
            blocks.set_fused_blocks(default_on)

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))

    a, b = runs
    assert rel(a[0], b[0]) < 1e-5 and rel(a[1], b[1]) < 1e-5
    assert a[2].keys() == b[2].keys() and len(a[2]) >= 7
    for k in a[2]:
        assert rel(a[2][k], b[2][k]) < 1e-5, k
    for k in a[3]:
        assert rel(a[3][k].float(), b[3][k].float()) < 1e-6, k
