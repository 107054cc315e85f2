"""Parity of the fused attention forward (sketch2img_b200/csrc/attn.cu, C ABI s2i_attention) against an fp64 torch
softmax(scale * Q K^T) V computed from the same fp16 operands.  Tolerance 3e-3 relative L2 (P is rounded to fp16
before the P V product, exactly like the unfused path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def L(cuda):
    from sketch2img_b200 import _lib
    _lib.lib()
    return _lib


def _run(L, q, ldq, q_c0, kv, ldkv, k_c0, v_c0, B, heads, Nq, Nk, dp, d, scale, want_lse=True):
    out = torch.full((B, Nq, heads * dp), float("nan"), device=q.device, dtype=torch.float16)
    lse = torch.full((B * heads, Nq), float("nan"), device=q.device) if want_lse else None
    L.check(L.lib().s2i_attention(q.data_ptr(), ldq, q_c0, kv.data_ptr(), ldkv, k_c0, v_c0, B, heads, Nq, Nk, dp, d, scale,
                                  out.data_ptr(), heads * dp, lse.data_ptr() if want_lse else None, L.stream_ptr()))
    torch.cuda.synchronize()
    return out, lse


@pytest.mark.parametrize("B,N,heads,d,dp", [(2, 256, 8, 40, 48), (1, 1024, 4, 80, 80), (2, 256, 8, 160, 160),
                                            (2, 4096, 2, 40, 48), (2, 256, 4, 16, 16), (1, 320, 5, 64, 64),
                                            (3, 576, 2, 64, 64), (2, 128, 8, 40, 48),
                                            # long, ragged key / query ranges with narrow heads
                                            (1, 1024, 2, 64, 64), (2, 1100, 3, 40, 48), (1, 1030, 1, 16, 16)])
def test_self_attention(L, cuda, B, N, heads, d, dp):
    """Q, K, V head-sliced out of one fused [B][N][3*heads*dp] projection (the UNet's qkv layout)."""
    g = torch.Generator(device="cpu").manual_seed(N + d)
    HP = heads * dp
    qkv = torch.randn(B, N, 3, heads, dp, generator=g)
    qkv[..., d:] = 0.0
    qkv = qkv.reshape(B, N, 3 * HP).to(cuda).half().contiguous()
    scale = d ** -0.5
    out, lse = _run(L, qkv, 3 * HP, 0, qkv, 3 * HP, HP, 2 * HP, B, heads, N, N, dp, d, scale)
    v5 = qkv.view(B, N, 3, heads, dp).double()
    q, k, v = (v5[:, :, i].permute(0, 2, 1, 3) for i in range(3))          # [B, h, N, dp]
    S = scale * q @ k.transpose(-1, -2)
    ref = (torch.softmax(S, -1) @ v).permute(0, 2, 1, 3).reshape(B, N, HP)
    assert rel(out.float(), ref) < 3e-3
    assert rel(lse, torch.logsumexp(S, -1).reshape(B * heads, N)) < 1e-4
    # a second launch on the same buffers (stale shared memory / TMEM from the previous CTA must not matter)
    out2, _ = _run(L, qkv, 3 * HP, 0, qkv, 3 * HP, HP, 2 * HP, B, heads, N, N, dp, d, scale, want_lse=False)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("Nq,Nk", [(256, 77), (4096, 77), (128, 64), (256, 200)])
def test_cross_attention(L, cuda, Nq, Nk):
    """K/V from a separate token tensor whose length is not a multiple of the key tile (text context: 77)."""
    B, heads, d, dp = 2, 8, 40, 48
    HP = heads * dp
    g = torch.Generator(device="cpu").manual_seed(Nq + Nk)
    q = torch.randn(B, Nq, heads, dp, generator=g)
    kv = torch.randn(B, Nk, 2, heads, dp, generator=g)
    q[..., d:] = 0.0
    kv[..., d:] = 0.0
    qd = q.reshape(B, Nq, HP).to(cuda).half().contiguous()
    kvd = kv.reshape(B, Nk, 2 * HP).to(cuda).half().contiguous()
    scale = 0.3
    out, lse = _run(L, qd, HP, 0, kvd, 2 * HP, 0, HP, B, heads, Nq, Nk, dp, d, scale)
    qq = qd.view(B, Nq, heads, dp).permute(0, 2, 1, 3).double()
    kk = kvd.view(B, Nk, 2, heads, dp)[:, :, 0].permute(0, 2, 1, 3).double()
    vv = kvd.view(B, Nk, 2, heads, dp)[:, :, 1].permute(0, 2, 1, 3).double()
    S = scale * qq @ kk.transpose(-1, -2)
    ref = (torch.softmax(S, -1) @ vv).permute(0, 2, 1, 3).reshape(B, Nq, HP)
    assert rel(out.float(), ref) < 3e-3
    assert rel(lse, torch.logsumexp(S, -1).reshape(B * heads, Nq)) < 1e-4


def test_large_logits_do_not_overflow(L, cuda):
    """Exact row maxima (pass 1) keep exp2 arguments <= 0 whatever the score scale."""
    B, N, heads, d, dp = 1, 256, 2, 64, 64
    HP = heads * dp
    g = torch.Generator(device="cpu").manual_seed(3)
    qkv = (torch.randn(B, N, 3 * HP, generator=g) * 6).to(cuda).half().contiguous()
    out, _ = _run(L, qkv, 3 * HP, 0, qkv, 3 * HP, HP, 2 * HP, B, heads, N, N, dp, d, 1.0, want_lse=False)
    v5 = qkv.view(B, N, 3, heads, dp).double()
    q, k, v = (v5[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).permute(0, 2, 1, 3).reshape(B, N, HP)
    assert torch.isfinite(out).all()
    assert rel(out.float(), ref) < 5e-3


def test_unsupported_shapes_are_rejected(L, cuda):
    x = torch.zeros(1, 32, 3 * 64, device=cuda, dtype=torch.float16)
    with pytest.raises(L.S2IError):
        _run(L, x, 192, 0, x, 192, 64, 128, 1, 1, 32, 32, 64, 64, 1.0)


# ------------------------------------------------------------------------------------------------ fused backward
def _bwd(L, q, ldq, q_c0, kv, ldkv, k_c0, v_c0, dO, out, lse, B, heads, Nq, Nk, dp, d, scale, self_attn):
    HP = heads * dp
    delta = torch.full((B * heads, Nq), float("nan"), device=q.device)
    dq = torch.full((B, Nq, HP), float("nan"), device=q.device, dtype=torch.float16)
    dkv = torch.full((B, Nk, 2 * HP), float("nan"), device=q.device, dtype=torch.float16) if self_attn else None
    L.check(L.lib().s2i_attention_backward(q.data_ptr(), ldq, q_c0, kv.data_ptr(), ldkv, k_c0, v_c0, dO.data_ptr(),
                                           out.data_ptr(), HP, lse.data_ptr(), delta.data_ptr(), B, heads, Nq, Nk, dp, d, scale,
                                           dq.data_ptr(), HP, 0, dkv.data_ptr() if self_attn else None, 2 * HP, 0, HP,
                                           L.stream_ptr()))
    torch.cuda.synchronize()
    return dq, dkv


@pytest.mark.parametrize("B,N,heads,d,dp", [(2, 256, 8, 40, 48), (1, 1024, 4, 80, 80), (2, 4096, 2, 40, 48),
                                            (1, 320, 5, 64, 64), (3, 576, 2, 64, 64), (2, 128, 8, 40, 48), (1, 200, 2, 16, 16),
                                            # SD1.5's 16 x 16 level: 160-wide heads (three 64-column operand chunks, dK | dV
                                            # accumulators in 320 TMEM columns next to single-buffered tile products)
                                            (2, 256, 8, 160, 160), (1, 300, 2, 160, 160), (1, 256, 2, 192, 192)])
def test_self_attention_backward(L, cuda, B, N, heads, d, dp):
    """dQ, dK, dV of softmax(scale Q K^T) V against fp64 autograd on the same fp16 operands (recompute from the
    forward's log-sum-exp; P and dS are rounded to fp16 before the accumulating products like the unfused path)."""
    g = torch.Generator(device="cpu").manual_seed(N * 3 + d)
    HP = heads * dp
    qkv = torch.randn(B, N, 3, heads, dp, generator=g)
    qkv[..., d:] = 0.0
    qkv = qkv.reshape(B, N, 3 * HP).to(cuda).half().contiguous()
    dO = torch.randn(B, N, heads, dp, generator=g)
    dO[..., d:] = 0.0
    dO = dO.reshape(B, N, HP).to(cuda).half().contiguous()
    scale = d ** -0.5
    out, lse = _run(L, qkv, 3 * HP, 0, qkv, 3 * HP, HP, 2 * HP, B, heads, N, N, dp, d, scale)
    dq, dkv = _bwd(L, qkv, 3 * HP, 0, qkv, 3 * HP, HP, 2 * HP, dO, out, lse, B, heads, N, N, dp, d, scale, True)
    v5 = qkv.view(B, N, 3, heads, dp).double()
    q, k, v = (v5[:, :, i].permute(0, 2, 1, 3).clone().requires_grad_(True) for i in range(3))
    o = torch.softmax(scale * q @ k.transpose(-1, -2), -1) @ v
    gq, gk, gv = torch.autograd.grad(o, (q, k, v), dO.view(B, N, heads, dp).permute(0, 2, 1, 3).double())
    back = lambda t: t.permute(0, 2, 1, 3).reshape(B, N, HP)
    assert rel(dq.float(), back(gq)) < 5e-3
    assert rel(dkv[:, :, :HP].float(), back(gk)) < 5e-3
    assert rel(dkv[:, :, HP:].float(), back(gv)) < 5e-3


@pytest.mark.parametrize("Nq,Nk,d,dp", [(256, 77, 40, 48), (4096, 77, 40, 48), (128, 64, 40, 48), (256, 200, 40, 48),
                                        (256, 77, 160, 160)])
def test_cross_attention_backward(L, cuda, Nq, Nk, d, dp):
    """Only dQ (the text context is a constant); ragged key count."""
    B, heads = 2, 8
    HP = heads * dp
    g = torch.Generator(device="cpu").manual_seed(Nq * 5 + Nk)
    q = torch.randn(B, Nq, heads, dp, generator=g)
    kv = torch.randn(B, Nk, 2, heads, dp, generator=g)
    dO = torch.randn(B, Nq, heads, dp, generator=g)
    for t in (q, kv, dO):
        t[..., d:] = 0.0
    qd = q.reshape(B, Nq, HP).to(cuda).half().contiguous()
    kvd = kv.reshape(B, Nk, 2 * HP).to(cuda).half().contiguous()
    dOd = dO.reshape(B, Nq, HP).to(cuda).half().contiguous()
    scale = 0.3
    out, lse = _run(L, qd, HP, 0, kvd, 2 * HP, 0, HP, B, heads, Nq, Nk, dp, d, scale)
    dq, _ = _bwd(L, qd, HP, 0, kvd, 2 * HP, 0, HP, dOd, out, lse, B, heads, Nq, Nk, dp, d, scale, False)
    qq = qd.view(B, Nq, heads, dp).permute(0, 2, 1, 3).double().requires_grad_(True)
    kk = kvd.view(B, Nk, 2, heads, dp)[:, :, 0].permute(0, 2, 1, 3).double()
    vv = kvd.view(B, Nk, 2, heads, dp)[:, :, 1].permute(0, 2, 1, 3).double()
    o = torch.softmax(scale * qq @ kk.transpose(-1, -2), -1) @ vv
    gq, = torch.autograd.grad(o, qq, dOd.view(B, Nq, heads, dp).permute(0, 2, 1, 3).double())
    assert rel(dq.float(), gq.permute(0, 2, 1, 3).reshape(B, Nq, HP)) < 5e-3
