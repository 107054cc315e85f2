"""Parity of the one-launch GroupNorm(32) [+ SiLU] kernels (sketch2img_b200/csrc/kernels.cu gn_cluster_kernel, C ABI
s2i_groupnorm_forward / s2i_groupnorm_backward) against torch.nn.functional.group_norm in fp64 on the same NHWC fp32
input -- the norm1/norm2 + nonlinearity of every ResnetBlock2D and the norm of every Transformer2DModel inside
`self.unet(...)` (modules/pipeline.py:96), forward and the autograd backward of :159.
Tolerances: forward 1e-3 relative L2 (the output is rounded to fp16: 2^-11 per element), fp32 backward 5e-5."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def L(cuda):
    from sketch2img_b200 import _lib
    _lib.lib()
    return _lib


# every (batch, tokens, width) combination the SD1.5 / SD2.1 UNets meet at 64x64 / 96x96 latents, B = 1 (cond-only
# backward), 2 (one CFG pair) and 8 (4 images per GPU), plus ragged pixel counts and a width whose group size is odd
SHAPES = [(2, 4096, 320), (2, 4096, 640), (2, 4096, 960), (2, 1024, 320), (2, 1024, 640), (2, 1024, 1280), (2, 1024, 1920),
          (2, 256, 1280), (2, 256, 2560), (2, 64, 1280), (2, 64, 2560), (1, 4096, 320), (1, 4096, 960), (1, 64, 2560),
          (8, 1024, 640), (3, 100, 96), (1, 9, 32), (2, 576, 352), (5, 9216, 320)]


def _ref(x, gamma, beta, eps, silu):
    B, HW, C = x.shape
    y = F.group_norm(x.double().permute(0, 2, 1), 32, gamma.double(), beta.double(), eps).permute(0, 2, 1)
    return F.silu(y) if silu else y


@pytest.mark.parametrize("B,HW,C", SHAPES)
@pytest.mark.parametrize("silu", [1, 0])
def test_groupnorm_forward_and_backward(L, cuda, B, HW, C, silu):
    g = torch.Generator().manual_seed(B * 1000 + HW + C)
    ld = C + 8                                           # a row stride wider than the tensor (concat buffers)
    xb = torch.randn(B, HW, ld, generator=g) * 1.7 + 0.4
    x = xb[:, :, :C]
    gamma = 1.0 + 0.2 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    dy = torch.randn(B, HW, C, generator=g)
    add = torch.randn(B, HW, C, generator=g)
    eps = 1e-5
    xd, gd, bd, dyd, addd = xb.to(cuda), gamma.to(cuda), beta.to(cuda), dy.to(cuda).contiguous(), add.to(cuda).contiguous()
    out = torch.full((B, HW, C), float("nan"), device=cuda, dtype=torch.float16)
    raw = torch.full((B, HW, C), float("nan"), device=cuda, dtype=torch.float16)
    st_f = torch.zeros(B * 64, device=cuda, dtype=torch.float64)
    st_b = torch.zeros(B * 64, device=cuda, dtype=torch.float64)
    lib = L.lib()
    L.check(lib.s2i_groupnorm_forward(xd.data_ptr(), ld, B, HW, C, gd.data_ptr(), bd.data_ptr(), eps, silu, out.data_ptr(), C,
                                      raw.data_ptr(), C, st_f.data_ptr(), L.stream_ptr()))
    torch.cuda.synchronize()
    xr = x.double().requires_grad_(True)
    with torch.enable_grad():
        want = _ref(xr, gamma, beta, eps, silu)
        dx_want = torch.autograd.grad((want * dy.double()).sum(), xr)[0] + add.double()
    assert rel(out.float().cpu(), want.detach()) < 1e-3
    assert torch.equal(raw.cpu(), x.half())
    # the statistics slot: (mean, rstd) per (sample, group) as float2
    stats = st_f.view(torch.float32).view(B, 128)[:, :64].reshape(B, 32, 2).cpu()
    xg = x.double().reshape(B, HW, 32, C // 32)
    assert torch.allclose(stats[..., 0].double(), xg.mean((1, 3)), atol=1e-5)
    assert torch.allclose(stats[..., 1].double(), (xg.var((1, 3), unbiased=False) + eps).rsqrt(), rtol=1e-5)
    dx32 = torch.full((B, HW, C), float("nan"), device=cuda)
    dx16 = torch.full((B, HW, C), float("nan"), device=cuda, dtype=torch.float16)
    L.check(lib.s2i_groupnorm_backward(dyd.data_ptr(), C, xd.data_ptr(), ld, B, HW, C, gd.data_ptr(), bd.data_ptr(), eps, silu,
                                       st_f.data_ptr(), st_b.data_ptr(), addd.data_ptr(), C, dx32.data_ptr(), C,
                                       dx16.data_ptr(), C, L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel(dx32.cpu(), dx_want) < 5e-5
    assert rel(dx16.float().cpu(), dx_want) < 1e-3
    # size-independent property: the backward is linear in dy
    st_b.zero_()
    dx2 = torch.empty_like(dx32)
    L.check(lib.s2i_groupnorm_backward((2 * dyd).data_ptr(), C, xd.data_ptr(), ld, B, HW, C, gd.data_ptr(), bd.data_ptr(), eps,
                                       silu, st_f.data_ptr(), st_b.data_ptr(), None, 0, dx2.data_ptr(), C, None, 0,
                                       L.stream_ptr()))
    torch.cuda.synchronize()
    assert rel(dx2.cpu(), 2 * (dx_want - add.double())) < 5e-5


def test_groupnorm_is_deterministic(L, cuda):
    """Partial sums are exchanged in rank order (no atomics): two launches give bit-identical results."""
    B, HW, C = 2, 4096, 320
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, HW, C, generator=g).to(cuda)
    gamma, beta = torch.ones(C, device=cuda), torch.zeros(C, device=cuda)
    outs = []
    for _ in range(2):
        out = torch.empty(B, HW, C, device=cuda, dtype=torch.float16)
        st = torch.zeros(B * 64, device=cuda, dtype=torch.float64)
        L.check(L.lib().s2i_groupnorm_forward(x.data_ptr(), C, B, HW, C, gamma.data_ptr(), beta.data_ptr(), 1e-5, 1,
                                              out.data_ptr(), C, None, 0, st.data_ptr(), L.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


def test_groupnorm_rejects_bad_arguments(L, cuda):
    x = torch.zeros(1, 16, 48, device=cuda)
    out = torch.zeros(1, 16, 48, device=cuda, dtype=torch.float16)
    st = torch.zeros(64, device=cuda, dtype=torch.float64)
    w = torch.ones(48, device=cuda)
    lib = L.lib()
    with pytest.raises(L.S2IError):      # 48 channels are not a multiple of 32 groups
        L.check(lib.s2i_groupnorm_forward(x.data_ptr(), 48, 1, 16, 48, w.data_ptr(), w.data_ptr(), 1e-5, 1, out.data_ptr(), 48,
                                          None, 0, st.data_ptr(), L.stream_ptr()))
    with pytest.raises(L.S2IError):      # null input
        L.check(lib.s2i_groupnorm_forward(None, 64, 1, 16, 64, w.data_ptr(), w.data_ptr(), 1e-5, 1, out.data_ptr(), 64,
                                          None, 0, st.data_ptr(), L.stream_ptr()))


@pytest.mark.parametrize("B,H,W,Cin,C,taps", [(2, 64, 64, 320, 320, 9), (2, 32, 32, 320, 640, 9), (2, 16, 16, 640, 1280, 1),
                                              (2, 8, 8, 1280, 1280, 9), (1, 64, 64, 320, 320, 1), (2, 32, 32, 960, 960, 1)])
@pytest.mark.parametrize("silu", [1, 0])
def test_groupnorm_forward_from_producer_statistics(L, cuda, B, H, W, Cin, C, taps, silu):
    """gn_norm_kernel (C ABI s2i_groupnorm_forward_colstat): the GEMM that produces x leaves per-block column sums
    (s2i_gemm_desc.colstat) and the GroupNorm takes its statistics from them -- ONE pass over x.  Same reference, same
    tolerance as the two-phase kernel, the per-group statistics handed to the backward included, and the same bits twice."""
    import ctypes as C_
    g = torch.Generator().manual_seed(B * 131 + H + C + taps)
    a = torch.randn(B, H, W, Cin, generator=g).to(cuda).half()
    w = (torch.randn(C, taps * Cin, generator=g) * 0.03).to(cuda).half()
    bias = (torch.randn(C, generator=g) * 0.5).to(cuda)
    gamma = (1.0 + 0.2 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(C, generator=g)).to(cuda)
    HW, eps = H * W, 1e-5
    cap = max(32, HW // 64)
    x = torch.zeros(B, H, W, C, device=cuda)
    stat = torch.zeros(B, cap, 2, C, device=cuda)
    bps = C_.c_int(0)
    d = L.GemmDesc(A=a.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=taps,
                   B=w.data_ptr(), bI=taps * Cin, bR=C, b_sr=taps * Cin, N=C, Kc=Cin, bias=bias.data_ptr(),
                   out32=x.data_ptr(), ld32=C, colstat=stat.data_ptr(), colstat_ld=C, colstat_cap=cap,
                   colstat_bps=C_.pointer(bps))
    L.gemm(d)
    torch.cuda.synchronize()
    assert bps.value > 0
    lib = L.lib()
    outs = []
    for rep in range(2):
        out = torch.full((B, HW, C), float("nan"), device=cuda, dtype=torch.float16)
        raw = torch.full((B, HW, C), float("nan"), device=cuda, dtype=torch.float16)
        slot = torch.zeros(B * 64, device=cuda, dtype=torch.float64)
        L.check(lib.s2i_groupnorm_forward_colstat(x.data_ptr(), C, B, HW, C, stat.data_ptr(), C, cap, bps.value, gamma.data_ptr(),
                                                  beta.data_ptr(), eps, silu, out.data_ptr(), C, raw.data_ptr(), C, slot.data_ptr(),
                                                  L.stream_ptr()))
        torch.cuda.synchronize()
        outs.append(out.clone())
    xf = x.reshape(B, HW, C)
    ref = _ref(xf.cpu(), gamma.cpu(), beta.cpu(), eps, silu)
    assert rel(outs[0].cpu(), ref) < 1e-3
    assert rel(raw.cpu().float(), xf.cpu()) < 1e-3
    assert torch.equal(outs[0], outs[1])
    # per-group mean / rstd in the slot (float2 per group): what the backward will read
    xg = xf.double().reshape(B, HW, 32, C // 32)
    mean = xg.mean(dim=(1, 3))
    rstd = 1.0 / torch.sqrt(xg.var(dim=(1, 3), unbiased=False) + eps)
    got = slot.view(torch.float32).reshape(B, 64, 2)[:, :32]
    assert (got[..., 0].double() - mean).abs().max() < 1e-4 * (1 + mean.abs().max())
    assert ((got[..., 1].double() - rstd) / rstd).abs().max() < 1e-4
