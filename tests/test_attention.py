"""Masked cross-attention of the Mask3D decoder (csrc/attention.cu).

CPU: the oracle's explicit restatement (oracle/ops_cpu.masked_attention_core) is pinned to what the reference executes —
nn.MultiheadAttention with a boolean memory_mask in the layout models/mask3d.py:355-362 builds.
GPU: the libus3d kernels, through the C ABI, against the oracle on the same seeded inputs, forward and all three
gradients; fp32 tolerance 2e-5 relative (north_star: 1e-3 on features).
"""
import pytest
import torch

from oracle import ops_cpu


def _case(Q, K, B, H, hd, seed, hidden=0.7):
    g = torch.Generator().manual_seed(seed)
    E = H * hd
    q = torch.randn(Q, B, E, generator=g)
    k = torch.randn(K, B, E, generator=g)
    v = torch.randn(K, B, E, generator=g)
    # the decoder's mask: [B, K, Q], True = hidden; rows that would hide everything are un-hidden (mask3d.py:349)
    m = torch.rand(B, K, Q, generator=g) < hidden
    m.permute(0, 2, 1)[m.sum(1) == K] = False
    return q, k, v, m


def _torch_layout(m_bkq, H):
    """models/mask3d.py:358: repeat_interleave over heads, permute to [B*h, Q, K]."""
    return m_bkq.repeat_interleave(H, dim=0).permute(0, 2, 1)


def test_oracle_attention_core_matches_nn_multihead_attention():
    Q, K, B, H, hd = 100, 333, 2, 8, 16
    q_in, k_in, v_in, m = _case(Q, K, B, H, hd, 0)
    torch.manual_seed(1)
    mha = torch.nn.MultiheadAttention(H * hd, H, dropout=0.0)
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
        mha.out_proj.bias.normal_(0, 0.1)
    ref = ops_cpu.multihead_cross_attention(mha, q_in, k_in, v_in, attn_mask=_torch_layout(m, H))
    E = H * hd
    w, b = mha.in_proj_weight, mha.in_proj_bias
    lin = torch.nn.functional.linear
    core = ops_cpu.masked_attention_core(lin(q_in, w[:E], b[:E]), lin(k_in, w[E:2 * E], b[E:2 * E]), lin(v_in, w[2 * E:], b[2 * E:]),
                                         m.permute(0, 2, 1)[:, None], H)
    got = mha.out_proj(core)
    assert float((got - ref).norm() / ref.norm()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("Q,K,B,H,hd,layout", [
    (100, 200, 2, 8, 16, "bkq"),      # hlevel 0 of the decoder (sample_sizes[0] = 200)
    (100, 3200, 4, 8, 16, "bkq"),     # C3 shape of hlevel 2
    (100, 1000, 2, 8, 16, "torch"),   # the reference's [B*h, Q, K] layout, ragged last key chunk
    (100, 130, 1, 8, 16, "none"),     # no mask
    (7, 1, 1, 2, 16, "bkq"),          # single key
    (150, 300, 2, 4, 32, "torch"),    # two query blocks, head_dim 32
])
def test_cuda_attention_core_matches_oracle(Q, K, B, H, hd, layout):
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.engine import functional as Fn

    q, k, v, m = _case(Q, K, B, H, hd, Q + K)
    if layout == "none":
        m = None
    qo, ko, vo = (t.clone().double().requires_grad_() for t in (q, k, v))
    ref = ops_cpu.masked_attention_core(qo, ko, vo, None if m is None else m.permute(0, 2, 1)[:, None], H)
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    ref.backward(g.double())

    qc, kc, vc = (t.clone().cuda().requires_grad_() for t in (q, k, v))
    if m is None:
        mask = None
    elif layout == "bkq":   # the decoder's own tensor, read in place
        mask = Fn.DecoderMask(m.cuda())
    else:
        mask = _torch_layout(m, H).cuda()
    out = Fn.MaskedCrossAttentionFunction.apply(qc, kc, vc, mask, H)
    out.backward(g.cuda())
    tol = 2e-5
    rel = lambda a, b: float((a.detach().double().cpu() - b.detach()).norm() / b.detach().norm().clamp(min=1e-30))
    assert rel(out, ref) < tol
    assert rel(qc.grad, qo.grad) < tol
    assert rel(kc.grad, ko.grad) < tol
    assert rel(vc.grad, vo.grad) < tol
    # no atomics anywhere: identical bits run to run
    qd, kd, vd = (t.clone().cuda().requires_grad_() for t in (q, k, v))
    out2 = Fn.MaskedCrossAttentionFunction.apply(qd, kd, vd, mask, H)
    out2.backward(g.cuda())
    assert torch.equal(out, out2) and torch.equal(qc.grad, qd.grad) and torch.equal(kc.grad, kd.grad) and torch.equal(vc.grad, vd.grad)


@pytest.mark.gpu
def test_cuda_multihead_cross_attention_matches_nn_module():
    """The layer-level entry point keeps nn.MultiheadAttention's parameters: same state dict, same result."""
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.engine import functional as Fn

    Q, K, B, H, hd = 100, 800, 3, 8, 16
    q, k, v, m = _case(Q, K, B, H, hd, 11)
    torch.manual_seed(2)
    mha = torch.nn.MultiheadAttention(H * hd, H, dropout=0.0)
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.1)
    ref_in = [t.clone().requires_grad_() for t in (q, k, v)]
    ref = ops_cpu.multihead_cross_attention(mha, *ref_in, attn_mask=_torch_layout(m, H))
    g = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    ref.backward(g)
    ref_wgrad = mha.in_proj_weight.grad.clone()
    mha.zero_grad()
    mha = mha.cuda()
    cu_in = [t.clone().cuda().requires_grad_() for t in (q, k, v)]
    for mask in (_torch_layout(m, H).cuda(), Fn.DecoderMask(m.cuda())):
        mha.zero_grad()
        for t in cu_in:
            t.grad = None
        out = Fn.multihead_cross_attention(mha, *cu_in, attn_mask=mask)
        out.backward(g.cuda())
        rel = lambda a, b: float((a.detach().cpu() - b.detach()).norm() / b.detach().norm())
        assert rel(out, ref) < 2e-5
        for a, b in zip(cu_in, ref_in):
            assert rel(a.grad, b.grad) < 2e-5
        assert rel(mha.in_proj_weight.grad, ref_wgrad) < 2e-5


@pytest.mark.gpu
def test_cuda_attention_rejects_what_it_does_not_implement():
    import unscene3d_b200  # noqa: F401
    from unscene3d_b200.engine import functional as Fn

    mha = torch.nn.MultiheadAttention(128, 8, dropout=0.1).cuda().train()
    x = torch.randn(4, 1, 128, device="cuda")
    with pytest.raises(RuntimeError, match="dropout"):
        Fn.multihead_cross_attention(mha, x, x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        Fn.MaskedCrossAttentionFunction.apply(x.cpu(), x.cpu(), x.cpu(), None, 8)
