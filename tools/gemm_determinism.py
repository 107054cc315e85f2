"""GPU box: bitwise run-to-run determinism of the tcgen05 GEMM on the shapes of the tiny UNet level 0."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib as L

def conv(B, H, W, Cin, Cout, reps=5):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().half()
    w = (torch.randn(Cout, 9 * Cin, generator=g) * 0.1).cuda().half()
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).cuda()
    outs = []
    for _ in range(reps):
        o = torch.full((B, H, W, Cout), float("nan"), device="cuda")
        d = L.GemmDesc(A=x.data_ptr(), aC=Cin, aW=W, aH=H, aB=B, a_sw=Cin, a_sh=Cin * W, a_sb=Cin * W * H, taps=9,
                       B=w.data_ptr(), bI=9 * Cin, bR=Cout, b_sr=9 * Cin, N=Cout, Kc=Cin, bias=bias.data_ptr(),
                       residual=res.data_ptr(), res_ld=Cout, out32=o.data_ptr(), ld32=Cout)
        L.gemm(d)
        torch.cuda.synchronize()
        outs.append(o.clone())
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float().reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2),
                                     bias, padding=1).permute(0, 2, 3, 1) + res
    print("conv %s: max |run_i - run_0| = %s ; rel err vs torch %.2e" % (
        (B, H, W, Cin, Cout), [float((o - outs[0]).abs().max()) for o in outs[1:]],
        float((outs[0] - ref).norm() / ref.norm())))

def linear(M, N, K, reps=5):
    g = torch.Generator().manual_seed(2)
    a = torch.randn(M, K, generator=g).cuda().half()
    w = torch.randn(N, K, generator=g).cuda().half()
    outs = []
    for _ in range(reps):
        o = torch.full((M, N), float("nan"), device="cuda")
        d = L.GemmDesc(A=a.data_ptr(), aC=K, aW=M, a_sw=K, B=w.data_ptr(), bI=K, bR=N, b_sr=K, N=N, Kc=K,
                       out32=o.data_ptr(), ld32=N)
        L.gemm(d)
        torch.cuda.synchronize()
        outs.append(o.clone())
    ref = a.float() @ w.float().t()
    print("linear %s: max |run_i - run_0| = %s ; rel err vs torch %.2e" % (
        (M, N, K), [float((o - outs[0]).abs().max()) for o in outs[1:]], float((outs[0] - ref).norm() / ref.norm())))

L.lib()
conv(4, 16, 16, 64, 64)
conv(2, 16, 16, 64, 64)
conv(4, 8, 8, 128, 128)
conv(2, 64, 64, 320, 320)
linear(1024, 64, 64)
linear(1024, 192, 64)
linear(1024, 512, 64)
linear(8192, 320, 320)
