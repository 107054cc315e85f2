"""GPU box: time the GEMM shapes of one SD1.5 CFG-pair step (B = 2) through s2i_gemm, with the TMA epilogue on and off,
optionally sweeping BN / split-K.  usage: python tools/gemm_bench.py [sweep]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib as L  # noqa: E402

# (name, pixels-per-side or None, rows, N, K, taps, residual, out32, out16)
SHAPES = [
    ("conv 320@64", 64, 8192, 320, 320, 9, 1, 1, 0),
    ("lin 320->320 @64 res", None, 8192, 320, 320, 1, 1, 1, 0),
    ("qkv 320->1152 @64", None, 8192, 1152, 320, 1, 0, 0, 1),
    ("ff1 320->2560 @64", None, 8192, 2560, 320, 1, 0, 0, 1),
    ("ff2 1280->320 @64", None, 8192, 320, 1280, 1, 1, 0, 1),
    ("conv 960->320@64", 64, 8192, 320, 960, 9, 1, 1, 0),
    ("conv 640@32", 32, 2048, 640, 640, 9, 1, 1, 0),
    ("lin 640->640 @32 res", None, 2048, 640, 640, 1, 1, 1, 0),
    ("ff1 640->5120 @32", None, 2048, 5120, 640, 1, 0, 0, 1),
    ("ff2 2560->640 @32", None, 2048, 640, 2560, 1, 1, 0, 1),
    ("conv 1280@16", 16, 512, 1280, 1280, 9, 1, 1, 0),
    ("conv 2560->1280@16", 16, 512, 1280, 2560, 9, 1, 1, 0),
    ("lin 1280->1280 @16 res", None, 512, 1280, 1280, 1, 1, 1, 0),
    ("ff1 1280->10240 @16", None, 512, 10240, 1280, 1, 0, 0, 1),
    ("ff2 5120->1280 @16", None, 512, 1280, 5120, 1, 1, 0, 1),
    ("conv 1280@8", 8, 128, 1280, 1280, 9, 1, 1, 0),
    ("conv 2560->1280@8", 8, 128, 1280, 2560, 9, 1, 1, 0),
    ("lin 1280->1280 @8 res", None, 128, 1280, 1280, 1, 1, 1, 0),
    # the cond-only input-gradient walk: one sample (half the rows)
    ("conv 320@64 B1", 64, 4096, 320, 320, 9, 1, 1, 0),
    ("lin 320->320 @64 B1", None, 4096, 320, 320, 1, 1, 1, 0),
    ("dff 2560->320 @64 B1", None, 4096, 320, 2560, 1, 0, 1, 0),
    ("conv 640@32 B1", 32, 1024, 640, 640, 9, 1, 1, 0),
    ("lin 640->640 @32 B1", None, 1024, 640, 640, 1, 1, 1, 0),
    ("conv 1280@16 B1", 16, 256, 1280, 1280, 9, 1, 1, 0),
    ("conv 1280@8 B1", 8, 64, 1280, 1280, 9, 1, 1, 0),
]


def make(shape):
    name, side, rows, N, K, taps, res, o32, o16 = shape
    g = torch.Generator().manual_seed(1)
    a = torch.randn(rows, K, generator=g).cuda().half()
    w = (torch.randn(N, taps * K, generator=g) * 0.02).cuda().half()
    bias = torch.randn(N, generator=g).cuda()
    r = torch.randn(rows, N, generator=g).cuda()
    out32 = torch.zeros(rows, N, device="cuda")
    out16 = torch.zeros(rows, N, device="cuda", dtype=torch.float16)
    kw = dict(A=a.data_ptr(), aC=K, B=w.data_ptr(), bI=taps * K, bR=N, b_sr=taps * K, N=N, Kc=K, taps=taps,
              bias=bias.data_ptr(), residual=r.data_ptr() if res else None, res_ld=N,
              out32=out32.data_ptr() if o32 else None, ld32=N, out16=out16.data_ptr() if o16 else None, ld16=N)
    if side:
        B = rows // (side * side)
        kw.update(aW=side, aH=side, aB=B, a_sw=K, a_sh=K * side, a_sb=K * side * side)
    else:
        kw.update(aW=rows, a_sw=K)
    return kw, (a, w, bias, r, out32, out16)


def time_desc(kw, reps=20, **over):
    """Device time per launch: `reps` launches captured in a CUDA graph (no host launch cost between them)."""
    d = L.GemmDesc(**{**kw, **over})
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            L.gemm(d)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                L.gemm(d)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def dsweep():
    """Deterministic split-K sweep: tile width x cluster size x groups per shape, against the cost model's own choice."""
    for shape in SHAPES:
        kw, keep = make(shape)
        name, side, rows, N, K, taps = shape[:6]
        os.environ.pop("S2I_GEMM_CS", None)
        t_def = time_desc(kw)
        res = []
        for bn in (64, 96, 128, 160, 192, 256):
            if bn > N:
                continue
            try:
                res.append((time_desc(kw, reps=10, BN=bn, splits=-1), bn, 1, 1))
            except L.S2IError:
                pass
            for cs in (1, 2, 4, 8, 16):
                for g in (1, 2, 3):
                    if cs * g == 1:
                        continue
                    os.environ["S2I_GEMM_CS"] = str(cs)
                    try:
                        res.append((time_desc(kw, reps=10, BN=bn, splits=cs * g), bn, cs, g))
                    except L.S2IError:
                        pass
                    os.environ.pop("S2I_GEMM_CS", None)
        res.sort()
        print(f"{name:28s} model's choice {t_def:6.1f} us; best (us, BN, cluster, groups): "
              + " ".join("(%.1f %d %d %d)" % r for r in res[:6]), flush=True)
        del keep


def pair():
    """CTA pairs (tcgen05.mma.cta_group::2, half a B tile per CTA) against single CTAs: same bits, device time per shape,
    at the model's tile width and at forced widths."""
    lib = L.lib()
    print(f"{'shape':28s} {'single us':>10s} {'pair us':>9s}   same bits   forced BN: (BN single pair)")
    for shape in SHAPES:
        kw, keep = make(shape)
        out32, out16 = keep[4], keep[5]
        lib.s2i_gemm_set_pair(0)
        L.gemm(L.GemmDesc(**kw))
        torch.cuda.synchronize()
        ref = (out32.clone(), out16.clone())
        t0 = time_desc(kw)
        out32.zero_(); out16.zero_()
        lib.s2i_gemm_set_pair(1)
        L.gemm(L.GemmDesc(**kw))
        torch.cuda.synchronize()
        same = torch.equal(ref[0], out32) and torch.equal(ref[1], out16)
        err = max((ref[0] - out32).abs().max().item(), (ref[1].float() - out16.float()).abs().max().item())
        t1 = time_desc(kw)
        lib.s2i_gemm_set_pair(-1)
        t2 = time_desc(kw)
        extra = ["auto %.1f  " % t2]
        for bn in (128, 192, 256):
            if bn > shape[3]:
                continue
            try:
                lib.s2i_gemm_set_pair(0)
                a = time_desc(kw, reps=10, BN=bn)
                lib.s2i_gemm_set_pair(1)
                b = time_desc(kw, reps=10, BN=bn)
                extra.append("(%d %.1f %.1f)" % (bn, a, b))
            except L.S2IError as e:
                extra.append("(%d err)" % bn)
        lib.s2i_gemm_set_pair(-1)
        print(f"{shape[0]:28s} {t0:10.1f} {t1:9.1f}   {str(same):5s} {err:.1e}   " + " ".join(extra), flush=True)
        del keep


def psweep():
    """Tile width x cluster split-K size, single CTAs against CTA pairs: best four of each per shape."""
    lib = L.lib()
    for shape in SHAPES:
        kw, keep = make(shape)
        line = f"{shape[0]:28s}"
        for pair_mode in (0, 1):
            lib.s2i_gemm_set_pair(pair_mode)
            res = []
            for bn in (64, 96, 128, 160, 192, 256):
                if bn > shape[3]:
                    continue
                for cs in (1, 2, 4, 8):
                    if cs > 1:
                        os.environ["S2I_GEMM_CS"] = str(cs)
                    try:
                        res.append((time_desc(kw, reps=10, BN=bn, splits=cs if cs > 1 else -1), bn, cs))
                    except L.S2IError:
                        pass
                    os.environ.pop("S2I_GEMM_CS", None)
            res.sort()
            line += ("  pair: " if pair_mode else "  single: ") + " ".join("(%.1f %d %d)" % r for r in res[:4])
        lib.s2i_gemm_set_pair(-1)
        print(line, flush=True)
        del keep


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "dsweep":
        return dsweep()
    if len(sys.argv) > 1 and sys.argv[1] == "psweep":
        return psweep()
    if len(sys.argv) > 1 and sys.argv[1] == "pair":
        return pair()
    sweep = len(sys.argv) > 1 and sys.argv[1] == "sweep"
    lib = L.lib()
    print(f"{'shape':28s} {'GFLOP':>7s} {'thread_epi us':>13s} {'tma_epi us':>11s} {'TF/s':>7s}")
    for shape in SHAPES:
        kw, keep = make(shape)
        name, side, rows, N, K, taps = shape[:6]
        gf = 2.0 * rows * N * K * taps / 1e9
        lib.s2i_gemm_set_tma_epilogue(0)
        t0 = time_desc(kw)
        lib.s2i_gemm_set_tma_epilogue(1)
        t1 = time_desc(kw)
        lib.s2i_gemm_force_msub(1)
        t1a = time_desc(kw)
        lib.s2i_gemm_force_msub(2)
        t1b = time_desc(kw)
        lib.s2i_gemm_force_msub(0)
        print(f"{name:28s} {gf:7.2f} {t0:13.1f} {t1:11.1f} {gf / t1:7.1f}   msub1 {t1a:6.1f}  msub2 {t1b:6.1f}", flush=True)
        if sweep:
            best = []
            for ms in (1, 2):
                lib.s2i_gemm_force_msub(ms)
                for bn in (64, 96, 128, 160, 192, 256):
                    if bn > N:
                        continue
                    for sp in (-1, 2, 4, 8, 16, 32):
                        if sp > 1 and shape[8]:
                            continue
                        try:
                            t = time_desc(kw, reps=10, BN=bn, splits=sp)
                        except L.S2IError:
                            continue
                        best.append((t, ms, bn, sp))
            lib.s2i_gemm_force_msub(0)
            best.sort()
            print("      best (us, msub, BN, splits):", [(round(t, 1), ms, bn, sp) for t, ms, bn, sp in best[:5]], flush=True)
        del keep


if __name__ == "__main__":
    main()
