"""Parity of the tcgen05/TMA implicit-GEMM kernel (sketch2img_b200/csrc/gemm_tc.cu) against fp64 torch
references computed from the same fp16/bf16-rounded operands.  Tolerance: 2e-3 relative L2 (fp32
accumulation-order noise only; operands are identical)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _dt(bf16):
    return torch.bfloat16 if bf16 else torch.float16


@pytest.fixture(scope="module", params=["tma_epilogue", "thread_epilogue"])
def L(cuda, request):
    """Every GEMM test runs twice: eligible shapes through gemm_tma_kernel (default), and everything through
    gemm_tc_kernel (the per-thread epilogue, still used for batched / MN-major / scaled / ReLU GEMMs)."""
    from sketch2img_b200 import _lib
    _lib.lib().s2i_gemm_set_tma_epilogue(1 if request.param == "tma_epilogue" else 0)
    _lib.tma_epilogue = request.param == "tma_epilogue"
    yield _lib
    _lib.lib().s2i_gemm_set_tma_epilogue(1)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 150, 200), (2048, 320, 320), (77, 96, 768), (8192, 16, 320)])
@pytest.mark.parametrize("bf16", [0, 1])
def test_linear(L, cuda, M, N, K, bf16):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    a = torch.randn(M, K, generator=g).to(cuda).to(_dt(bf16))
    w = torch.randn(N, K, generator=g).to(cuda).to(_dt(bf16))
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    out32 = torch.zeros(M, N, device=cuda)
    out16 = torch.zeros(M, N, device=cuda, dtype=torch.float16)
    d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=K, B=w.data_ptr(), bI=K, bR=N, b_sr=K, N=N, Kc=K, bf16=bf16,
                   alpha=0.5, bias=bias.data_ptr(), residual=res.data_ptr(), res_ld=N,
                   out32=out32.data_ptr(), ld32=N, out16=out16.data_ptr(), ld16=N)
    L.gemm(d)
    torch.cuda.synchronize()
    ref = 0.5 * (a.double() @ w.double().t()) + bias.double() + res.double()
    assert rel(out32, ref) < 2e-3
    assert rel(out16.float(), ref) < 3e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 16, 16, 64, 32), (2, 8, 8, 128, 48), (1, 64, 64, 64, 64),
                                            (2, 4, 4, 128, 64), (2, 2, 2, 64, 32), (3, 32, 32, 192, 80),
                                            (2, 24, 24, 64, 32)])
def test_conv3x3(L, cuda, B, H, W, Cin, Cout):
    g = torch.Generator(device="cpu").manual_seed(B * 131 + H)
    x = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()           # NHWC
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.1).to(cuda).half()
    bias = torch.randn(Cout, generator=g).to(cuda)
    temb = torch.randn(B, Cout, generator=g).to(cuda)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()       # [Cout][tap][Cin]
    out32 = torch.zeros(B, H, W, Cout, device=cuda)
    d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=9,
                   B=wp.data_ptr(), bI=9 * Cin, bR=Cout, b_sr=9 * Cin, N=Cout, Kc=Cin, bias=bias.data_ptr(),
                   rowvec=temb.data_ptr(), rowvec_ld=Cout, relu=1, out32=out32.data_ptr(), ld32=Cout)
    L.gemm(d)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=1)
    ref = torch.relu(ref + temb.double()[:, :, None, None]).permute(0, 2, 3, 1)
    assert rel(out32, ref) < 2e-3


def _attn_inputs(cuda, B, N, heads, d, dp, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    HP = heads * dp
    qkv = torch.randn(B, N, 3, heads, dp, generator=g)
    qkv[..., d:] = 0.0                      # head padding is zero in the real layout
    return qkv.reshape(B, N, 3 * HP).to(cuda).half().contiguous(), HP


@pytest.mark.parametrize("B,N,heads,d,dp", [(2, 256, 8, 40, 48), (2, 64, 8, 160, 160), (1, 1024, 4, 80, 80),
                                            (2, 16, 4, 16, 16), (2, 4096, 2, 40, 48)])
def test_attention_products(L, cuda, B, N, heads, d, dp):
    qkv, HP = _attn_inputs(cuda, B, N, heads, d, dp, N + d)
    Z = B * heads
    ldS = (N + 3) // 4 * 4
    ldP = (N + 7) // 8 * 8
    scale = d ** -0.5
    q = qkv.view(B, N, 3, heads, dp)[:, :, 0].permute(0, 2, 1, 3).reshape(Z, N, dp).double()
    k = qkv.view(B, N, 3, heads, dp)[:, :, 1].permute(0, 2, 1, 3).reshape(Z, N, dp).double()
    v = qkv.view(B, N, 3, heads, dp)[:, :, 2].permute(0, 2, 1, 3).reshape(Z, N, dp).double()

    # S = scale * Q K^T  (K-major x K-major, head-sliced operands, K = dp not a multiple of 64)
    S = torch.zeros(Z, N, ldS, device=cuda)
    d1 = L.GemmDesc(A=qkv.data_ptr(), aC=3 * HP, aW=N, aB=B, a_sw=3 * HP, a_sb=N * 3 * HP, a_hoff=dp,
                    B=qkv.data_ptr(), bI=3 * HP, bR=N, bZ=B, b_sr=3 * HP, b_sz=N * 3 * HP, b_c0=HP, b_hoff=dp,
                    N=N, Kc=dp, Z=Z, zh=heads, alpha=scale, out32=S.data_ptr(), ld32=ldS,
                    c_sb=heads * N * ldS, c_sh=N * ldS)
    L.gemm(d1)
    torch.cuda.synchronize()
    Sref = scale * q @ k.transpose(1, 2)
    assert rel(S[:, :, :N], Sref) < 2e-3

    # O = P V  (K-major P with batch = z, MN-major V sliced per head)
    P = torch.softmax(Sref, -1).half()
    Pbuf = torch.zeros(Z, N, ldP, device=cuda, dtype=torch.float16)
    Pbuf[:, :, :N] = P
    O = torch.zeros(B, N, HP, device=cuda, dtype=torch.float16)
    d2 = L.GemmDesc(A=Pbuf.data_ptr(), aC=N, aW=N, aB=Z, a_sw=ldP, a_sb=N * ldP, a_zmode=1,
                    B=qkv.data_ptr(), b_mn=1, bI=3 * HP, bR=N, bZ=B, b_sr=3 * HP, b_sz=N * 3 * HP, b_c0=2 * HP,
                    b_hoff=dp, N=dp, BN=dp, Kc=N, Z=Z, zh=heads, out16=O.data_ptr(), ld16=HP, c_sb=N * HP, c_sh=dp)
    L.gemm(d2)
    torch.cuda.synchronize()
    Oref = (P.double() @ v).reshape(B, heads, N, dp).permute(0, 2, 1, 3).reshape(B, N, HP)
    assert rel(O.float(), Oref) < 3e-3

    # dV = P^T dO  (MN-major A and MN-major B)
    dO = torch.randn(B, N, HP, device=cuda).half()
    dV = torch.zeros(B, N, 3 * HP, device=cuda, dtype=torch.float16)
    d3 = L.GemmDesc(A=Pbuf.data_ptr(), a_mn=1, aC=N, aW=N, aB=Z, a_sw=ldP, a_sb=N * ldP, a_zmode=1,
                    B=dO.data_ptr(), b_mn=1, bI=HP, bR=N, bZ=B, b_sr=HP, b_sz=N * HP, b_hoff=dp,
                    N=dp, BN=dp, Kc=N, Z=Z, zh=heads, out16=dV.data_ptr() + 2 * HP * 2, ld16=3 * HP,
                    c_sb=N * 3 * HP, c_sh=dp)
    L.gemm(d3)
    torch.cuda.synchronize()
    dOh = dO.view(B, N, heads, dp).permute(0, 2, 1, 3).reshape(Z, N, dp).double()
    dVref = (P.double().transpose(1, 2) @ dOh).reshape(B, heads, N, dp).permute(0, 2, 1, 3).reshape(B, N, HP)
    assert rel(dV[:, :, 2 * HP:].float(), dVref) < 3e-3

    # dP = dO V^T (fp32 out) and dQ = dS K (MN-major B = K section)
    dP = torch.zeros(Z, N, ldS, device=cuda)
    d4 = L.GemmDesc(A=dO.data_ptr(), aC=HP, aW=N, aB=B, a_sw=HP, a_sb=N * HP, a_hoff=dp,
                    B=qkv.data_ptr(), bI=3 * HP, bR=N, bZ=B, b_sr=3 * HP, b_sz=N * 3 * HP, b_c0=2 * HP, b_hoff=dp,
                    N=N, Kc=dp, Z=Z, zh=heads, out32=dP.data_ptr(), ld32=ldS, c_sb=heads * N * ldS, c_sh=N * ldS)
    L.gemm(d4)
    torch.cuda.synchronize()
    assert rel(dP[:, :, :N], dOh @ v.transpose(1, 2)) < 2e-3


def test_cross_attention_shapes(L, cuda):
    """Nk = 77 (text tokens): ragged K extent for PV and N extent for QK^T."""
    B, N, heads, dp, Nk = 2, 256, 8, 48, 77
    HP = heads * dp
    g = torch.Generator(device="cpu").manual_seed(5)
    q = torch.randn(B, N, HP, generator=g).to(cuda).half()
    kv = torch.randn(B, Nk, 2 * HP, generator=g).to(cuda).half()
    Z = B * heads
    ldS, ldP = 80, 80
    S = torch.full((Z, N, ldS), float("nan"), device=cuda)
    d1 = L.GemmDesc(A=q.data_ptr(), aC=HP, aW=N, aB=B, a_sw=HP, a_sb=N * HP, a_hoff=dp,
                    B=kv.data_ptr(), bI=2 * HP, bR=Nk, bZ=B, b_sr=2 * HP, b_sz=Nk * 2 * HP, b_hoff=dp,
                    N=Nk, Kc=dp, Z=Z, zh=heads, out32=S.data_ptr(), ld32=ldS, c_sb=heads * N * ldS, c_sh=N * ldS)
    L.gemm(d1)
    torch.cuda.synchronize()
    qh = q.view(B, N, heads, dp).permute(0, 2, 1, 3).reshape(Z, N, dp).double()
    kh = kv[:, :, :HP].reshape(B, Nk, heads, dp).permute(0, 2, 1, 3).reshape(Z, Nk, dp).double()
    vh = kv[:, :, HP:].reshape(B, Nk, heads, dp).permute(0, 2, 1, 3).reshape(Z, Nk, dp).double()
    Sref = qh @ kh.transpose(1, 2)
    assert rel(S[:, :, :Nk], Sref) < 2e-3
    P = torch.softmax(Sref * 0.1, -1).half()
    Pbuf = torch.full((Z, N, ldP), float("nan"), device=cuda, dtype=torch.float16)   # pad must never be read
    Pbuf[:, :, :Nk] = P
    O = torch.zeros(B, N, HP, device=cuda, dtype=torch.float16)
    d2 = L.GemmDesc(A=Pbuf.data_ptr(), aC=Nk, aW=N, aB=Z, a_sw=ldP, a_sb=N * ldP, a_zmode=1,
                    B=kv.data_ptr(), b_mn=1, bI=2 * HP, bR=Nk, bZ=B, b_sr=2 * HP, b_sz=Nk * 2 * HP, b_c0=HP, b_hoff=dp,
                    N=dp, BN=dp, Kc=Nk, Z=Z, zh=heads, out16=O.data_ptr(), ld16=HP, c_sb=N * HP, c_sh=dp)
    L.gemm(d2)
    torch.cuda.synchronize()
    Oref = (P.double() @ vh).reshape(B, heads, N, dp).permute(0, 2, 1, 3).reshape(B, N, HP)
    assert rel(O.float(), Oref) < 3e-3


@pytest.mark.parametrize("M,N,K,splits,BN", [(128, 1280, 2048, 8, 128), (512, 1280, 1280, 0, 0), (128, 320, 4096, 16, 64),
                                             (100, 64, 1024, 4, 0), (128, 10240, 1280, 0, 0), (2048, 640, 640, 0, 0)])
def test_split_k_linear(L, cuda, M, N, K, splits, BN):
    """Small-M GEMMs (the 8x8 / 16x16 levels) split K across CTAs; the last CTA per tile reduces the partial tiles in
    split order and applies the full epilogue (bias, residual, fp32 + fp16 stores)."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda).half()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda).half()
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).to(cuda)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    outs = []
    for rep in range(3):            # the arrival counters reset themselves: repeated launches stay correct
        out32 = torch.full((M, N), float("nan"), device=cuda)
        out16 = torch.zeros(M, N, device=cuda, dtype=torch.float16)
        d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=K, B=w.data_ptr(), bI=K, bR=N, b_sr=K, N=N, Kc=K, BN=BN,
                       splits=splits, bias=bias.data_ptr(), residual=res.data_ptr(), res_ld=N,
                       out32=out32.data_ptr(), ld32=N, out16=out16.data_ptr(), ld16=N)
        L.gemm(d)
        torch.cuda.synchronize()
        assert rel(out32, ref) < 2e-3
        assert rel(out16.float(), ref) < 3e-3
        outs.append(out32.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])     # reduction order is fixed


@pytest.mark.parametrize("M,N,K,splits,BN,res,o32,o16", [
    (8192, 320, 320, 0, 0, 1, 1, 0), (8192, 1152, 320, 0, 0, 0, 0, 1), (2048, 640, 2560, 0, 0, 1, 0, 1),
    (300, 96, 200, 0, 0, 1, 1, 1), (154, 768, 768, 0, 0, 0, 0, 1), (512, 1280, 1280, 4, 0, 1, 1, 0),
    (128, 1280, 5120, 0, 0, 1, 1, 0), (128, 320, 4096, 16, 64, 0, 1, 0), (8192, 2560, 320, 0, 256, 0, 1, 0),
    (1000, 32, 64, 0, 0, 1, 1, 1), (4096, 640, 1920, 0, 160, 1, 1, 1)])
def test_linear_plain_epilogue(L, cuda, M, N, K, splits, BN, res, o32, o16):
    """alpha = 1, bias + shared column vector (+ fp32 residual), fp32 and/or fp16 outputs: the shapes of the sampling
    path that take the TMA epilogue (ragged M, ragged last N tile, fp16-only output, split-K by reduce-add)."""
    g = torch.Generator(device="cpu").manual_seed(M + 3 * N + K)
    a = torch.randn(M, K, generator=g).to(cuda).half()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda).half()
    bias = torch.randn(N, generator=g).to(cuda)
    vec = torch.randn(N, generator=g).to(cuda)
    r = torch.randn(M, N, generator=g).to(cuda)
    ref = a.double() @ w.double().t() + bias.double() + vec.double() + (r.double() if res else 0.0)
    for rep in range(2):
        out32 = torch.full((M, N), float("nan"), device=cuda)
        out16 = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float16)
        d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=K, B=w.data_ptr(), bI=K, bR=N, b_sr=K, N=N, Kc=K, BN=BN,
                       splits=splits, bias=bias.data_ptr(), rowvec=vec.data_ptr(), rowvec_ld=0,
                       residual=r.data_ptr() if res else None, res_ld=N,
                       out32=out32.data_ptr() if o32 else None, ld32=N, out16=out16.data_ptr() if o16 else None, ld16=N)
        L.gemm(d)
        torch.cuda.synchronize()
        if o32:
            assert rel(out32, ref) < 2e-3
        if o16:
            assert rel(out16.float(), ref) < 3e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                            (2, 8, 8, 1280, 1280), (2, 24, 24, 64, 96), (3, 4, 4, 128, 64), (2, 2, 2, 64, 32)])
def test_conv3x3_plain_epilogue(L, cuda, B, H, W, Cin, Cout):
    """ResBlock conv with bias + time-embedding column vector + fp32 residual into a strided (wider) output."""
    g = torch.Generator(device="cpu").manual_seed(B * 17 + Cin + H)
    x = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.02).to(cuda).half()
    bias = torch.randn(Cout, generator=g).to(cuda)
    temb = torch.randn(Cout, generator=g).to(cuda)
    r = torch.randn(B, H, W, Cout + 32, generator=g).to(cuda)              # residual rows wider than N
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out32 = torch.full((B, H, W, Cout + 64), float("nan"), device=cuda)    # output rows wider than N
    d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=9,
                   B=wp.data_ptr(), bI=9 * Cin, bR=Cout, b_sr=9 * Cin, N=Cout, Kc=Cin, bias=bias.data_ptr(),
                   rowvec=temb.data_ptr(), rowvec_ld=0, residual=r.data_ptr(), res_ld=Cout + 32,
                   out32=out32.data_ptr(), ld32=Cout + 64)
    L.gemm(d)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double() + temb.double(), padding=1).permute(0, 2, 3, 1)
    ref = ref + r[..., :Cout].double()
    assert rel(out32[..., :Cout], ref) < 2e-3
    assert torch.isnan(out32[..., Cout:]).all()                            # nothing written outside the N columns


@pytest.mark.parametrize("B,H,W,Cin,Cout,splits", [(2, 8, 8, 1280, 1280, 0), (2, 16, 16, 640, 1280, 0), (2, 8, 8, 256, 64, 6)])
def test_split_k_conv3x3(L, cuda, B, H, W, Cin, Cout, splits):
    g = torch.Generator(device="cpu").manual_seed(B * 7 + Cin)
    x = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.02).to(cuda).half()
    bias = torch.randn(Cout, generator=g).to(cuda)
    temb = torch.randn(Cout, generator=g).to(cuda)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out32 = torch.full((B, H, W, Cout), float("nan"), device=cuda)
    d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=9,
                   B=wp.data_ptr(), bI=9 * Cin, bR=Cout, b_sr=9 * Cin, N=Cout, Kc=Cin, splits=splits, bias=bias.data_ptr(),
                   rowvec=temb.data_ptr(), rowvec_ld=0, out32=out32.data_ptr(), ld32=Cout)
    L.gemm(d)
    torch.cuda.synchronize()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double() + temb.double(), padding=1).permute(0, 2, 3, 1)
    assert rel(out32, ref) < 2e-3


@pytest.mark.parametrize("M,N,K,res,o32,o16,splits", [
    (8192, 320, 320, 1, 1, 0, 0), (300, 96, 200, 1, 1, 1, 0), (2048, 640, 2560, 1, 0, 1, 0), (512, 1280, 1280, 1, 1, 0, 4),
    (8192, 1152, 320, 0, 0, 1, 0), (384, 256, 512, 0, 1, 0, 0), (1000, 512, 4096, 1, 1, 0, 2)])
def test_two_subtile_linear(L, cuda, M, N, K, res, o32, o16, splits):
    """gemm_tma_kernel with two 128-row A tiles per CTA sharing each B tile (two TMEM accumulators, staging reused between
    the sub-tiles): ragged M (odd tile counts, partial last tile), residual reload in the second phase, split-K."""
    if not L.tma_epilogue:
        pytest.skip("two-sub-tile form belongs to the TMA-epilogue kernel")
    g = torch.Generator(device="cpu").manual_seed(M + 5 * N + K)
    a = torch.randn(M, K, generator=g).to(cuda).half()
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda).half()
    bias = torch.randn(N, generator=g).to(cuda)
    r = torch.randn(M, N, generator=g).to(cuda)
    ref = a.double() @ w.double().t() + bias.double() + (r.double() if res else 0.0)
    L.lib().s2i_gemm_force_msub(2)
    try:
        out32 = torch.full((M, N), float("nan"), device=cuda)
        out16 = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float16)
        d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=K, B=w.data_ptr(), bI=K, bR=N, b_sr=K, N=N, Kc=K, splits=splits,
                       bias=bias.data_ptr(), residual=r.data_ptr() if res else None, res_ld=N,
                       out32=out32.data_ptr() if o32 else None, ld32=N, out16=out16.data_ptr() if o16 else None, ld16=N)
        L.gemm(d)
        torch.cuda.synchronize()
    finally:
        L.lib().s2i_gemm_force_msub(0)
    if o32:
        assert rel(out32, ref) < 2e-3
    if o16:
        assert rel(out16.float(), ref) < 3e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 16, 16, 1280, 1280), (3, 8, 8, 256, 64), (2, 24, 24, 64, 96)])
def test_two_subtile_conv3x3(L, cuda, B, H, W, Cin, Cout):
    if not L.tma_epilogue:
        pytest.skip("two-sub-tile form belongs to the TMA-epilogue kernel")
    g = torch.Generator(device="cpu").manual_seed(B * 19 + Cin + H)
    x = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.02).to(cuda).half()
    bias = torch.randn(Cout, generator=g).to(cuda)
    r = torch.randn(B, H, W, Cout, generator=g).to(cuda)
    wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out32 = torch.full((B, H, W, Cout), float("nan"), device=cuda)
    L.lib().s2i_gemm_force_msub(2)
    try:
        d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=9,
                       B=wp.data_ptr(), bI=9 * Cin, bR=Cout, b_sr=9 * Cin, N=Cout, Kc=Cin, bias=bias.data_ptr(),
                       residual=r.data_ptr(), res_ld=Cout, out32=out32.data_ptr(), ld32=Cout)
        L.gemm(d)
        torch.cuda.synchronize()
    finally:
        L.lib().s2i_gemm_force_msub(0)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1) + r.double()
    assert rel(out32, ref) < 2e-3


@pytest.mark.parametrize("M,N,K", [(8192, 512, 9320), (8192, 256, 512), (1000, 64, 128)])
def test_relu_and_fp16_rounding_epilogues(L, cuda, M, N, K):
    """The LGP MLP's epilogues: ReLU on the forward Linear, and the backward's emulation of unscaled fp16 autograd
    rounding  v -> fp16(v / q) * q  (both also on the TMA-epilogue kernel)."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    a = torch.zeros(M, Kp)
    a[:, :K] = torch.randn(M, K, generator=g)
    a = a.to(cuda).half()
    w = torch.zeros(N, Kp)
    w[:, :K] = torch.randn(N, K, generator=g) * 0.02
    w = w.to(cuda).half()
    bias = torch.randn(N, generator=g).to(cuda)
    ref = a.double() @ w.double().t() + bias.double()
    out16 = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float16)
    d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=Kp, B=w.data_ptr(), bI=K, bR=N, b_sr=Kp, N=N, Kc=K, bias=bias.data_ptr(),
                   relu=1, out16=out16.data_ptr(), ld16=N)
    L.gemm(d)
    torch.cuda.synchronize()
    assert rel(out16.float(), torch.relu(ref)) < 3e-3
    q = 2.0 ** 14
    out32 = torch.full((M, N), float("nan"), device=cuda)
    d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=Kp, B=w.data_ptr(), bI=K, bR=N, b_sr=Kp, N=N, Kc=K, qscale=q,
                   out32=out32.data_ptr(), ld32=N)
    L.gemm(d)
    torch.cuda.synchronize()
    want = ((a.double() @ w.double().t()) / q).half().double() * q
    # identical up to the rare value whose fp32 accumulation-order noise crosses an fp16 rounding boundary
    assert rel(out32, want) < 2e-3
    assert torch.equal(out32, (out32 / q).half().float() * q)              # every output is an fp16 multiple of q


@pytest.mark.parametrize("M,C,keep,msub", [(8192, 320, True, 0), (8192, 320, False, 2), (2048, 640, True, 1), (512, 1280, False, 0),
                                           (300, 64, True, 0)])
def test_gated_gelu_epilogue(L, cuda, M, C, keep, msub):
    """The GEGLU feed-forward projection of every BasicTransformerBlock (diffusers FeedForward / GEGLU inside
    modules/pipeline.py:96):  out = value * gelu(gate),  [value | gate] = x W^T + b,  with the weight rows interleaved in blocks
    of 32 (value features 32 b .. at packed rows [64 b, 64 b + 32), their gates at [64 b + 32, 64 b + 64)) so the GEMM's
    epilogue applies the gate; `keep` also writes the interleaved projection (the backward's input).  Against fp64 torch."""
    import torch.nn.functional as F
    if not L.tma_epilogue:
        pytest.skip("the gated-GELU epilogue only exists in gemm_tma_kernel")
    g = torch.Generator(device="cpu").manual_seed(M + C)
    N, Fw = 8 * C, 4 * C
    x = torch.randn(M, C, generator=g).to(cuda).half()
    w = (torch.randn(N, C, generator=g) * (C ** -0.5))
    b = torch.randn(N, generator=g) * 0.1
    ref_ff = x.double().cpu() @ w.half().double().t() + b.double()
    want = ref_ff[:, :Fw] * F.gelu(ref_ff[:, Fw:])
    perm = torch.tensor([(r // 64) * 32 + r % 64 if r % 64 < 32 else Fw + (r // 64) * 32 + (r % 64 - 32) for r in range(N)])
    wp, bp = w[perm].to(cuda).half().contiguous(), b[perm].to(cuda).contiguous()
    out = torch.full((M, Fw), float("nan"), device=cuda, dtype=torch.float16)
    ff = torch.full((M, N), float("nan"), device=cuda, dtype=torch.float16)
    L.lib().s2i_gemm_force_msub(msub)
    try:
        d = L.GemmDesc(A=x.data_ptr(), aC=C, aW=M, a_sw=C, B=wp.data_ptr(), bI=C, bR=N, b_sr=C, N=N, Kc=C, bias=bp.data_ptr(),
                       out_glu=out.data_ptr(), ld_glu=Fw, out16=ff.data_ptr() if keep else None, ld16=N if keep else 0)
        L.gemm(d)
        torch.cuda.synchronize()
    finally:
        L.lib().s2i_gemm_force_msub(0)
    assert rel(out.float().cpu(), want) < 2e-3
    if keep:
        assert rel(ff.float().cpu(), ref_ff[:, perm]) < 1e-3
    # a tile width that cannot hold whole value / gate pairs is an error, as are the epilogue forms it does not combine with
    with pytest.raises(L.S2IError):
        L.gemm(L.GemmDesc(A=x.data_ptr(), aC=C, aW=M, a_sw=C, B=wp.data_ptr(), bI=C, bR=N, b_sr=C, N=N, Kc=C, BN=96,
                          out_glu=out.data_ptr(), ld_glu=Fw))
    o32 = torch.empty(M, N, device=cuda)
    with pytest.raises(L.S2IError):
        L.gemm(L.GemmDesc(A=x.data_ptr(), aC=C, aW=M, a_sw=C, B=wp.data_ptr(), bI=C, bR=N, b_sr=C, N=N, Kc=C,
                          out_glu=out.data_ptr(), ld_glu=Fw, out32=o32.data_ptr(), ld32=N))


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [(2, 64, 64, 320, 320, 9), (2, 32, 32, 640, 640, 9), (2, 16, 16, 1280, 1280, 9),
                                                 (2, 8, 8, 1280, 1280, 9), (2, 64, 64, 320, 320, 1), (2, 16, 16, 1280, 1280, 1),
                                                 (1, 8, 8, 1280, 640, 9), (2, 32, 32, 960, 640, 9)])
def test_column_statistics_for_groupnorm(L, cuda, B, H, W, Cin, Cout, taps):
    """GemmDesc.colstat: the per-block column sums / sums of squares a following GroupNorm takes its statistics from
    (norm1 / norm2 of diffusers' ResnetBlock2D inside modules/pipeline.py:96).  Summed over the blocks of a sample they must equal
    the sums of the fp32 result the same launch wrote, whichever form the launch takes (single CTAs, CTA pairs, cluster split-K),
    and be the same bits on every run."""
    if not L.tma_epilogue:
        pytest.skip("column statistics come from the TMA-epilogue kernel")
    import ctypes as C
    g = torch.Generator(device="cpu").manual_seed(B * 31 + Cin + H + taps)
    x = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()
    w = (torch.randn(Cout, taps * Cin, generator=g) * 0.02).to(cuda).half()
    bias = torch.randn(Cout, generator=g).to(cuda)
    r = torch.randn(B, H, W, Cout, generator=g).to(cuda)
    cap = max(32, H * W // 64)
    runs = []
    for rep in range(2):
        out32 = torch.zeros(B, H, W, Cout, device=cuda)
        stat = torch.full((B, cap, 2, Cout), float("nan"), device=cuda)
        bps = C.c_int(-1)
        d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=taps,
                       B=w.data_ptr(), bI=taps * Cin, bR=Cout, b_sr=taps * Cin, N=Cout, Kc=Cin, bias=bias.data_ptr(),
                       residual=r.data_ptr(), res_ld=Cout, out32=out32.data_ptr(), ld32=Cout,
                       colstat=stat.data_ptr(), colstat_ld=Cout, colstat_cap=cap, colstat_bps=C.pointer(bps))
        L.gemm(d)
        torch.cuda.synchronize()
        n = bps.value
        assert 0 < n <= cap, f"the launch did not provide statistics (bps = {n})"
        used = stat[:, :n].double()
        assert torch.isfinite(used).all()
        assert torch.isnan(stat[:, n:]).all()                               # nothing written beyond the blocks in use
        flat = out32.double().reshape(B, H * W, Cout)
        s1, s2 = flat.sum(1), (flat * flat).sum(1)
        assert (used[:, :, 0].sum(1) - s1).abs().max() <= 1e-4 * s1.abs().max() + 1e-2
        assert ((used[:, :, 1].sum(1) - s2).abs() / s2).max() < 1e-5
        runs.append(stat[:, :n].clone())
    assert torch.equal(runs[0], runs[1])
