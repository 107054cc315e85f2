"""GPU box: per-shape device time of the one-launch GroupNorm (+SiLU) forward / backward at the UNet's shapes.
usage: python tools/gn_bench.py            (S2I_GN_CLUSTER=0 selects the grid-barrier form for an A/B)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib  # noqa: E402

SHAPES = [(2, 4096, 320), (2, 4096, 640), (2, 4096, 960), (2, 1024, 640), (2, 1024, 1280), (2, 1024, 1920), (2, 256, 1280),
          (2, 256, 2560), (2, 64, 1280), (2, 64, 2560), (1, 4096, 320), (1, 4096, 960), (1, 1024, 1280), (1, 256, 2560),
          (8, 4096, 320), (8, 1024, 1280)]


def time_shape(lib, dev, B, HW, C, n=50):
    x = torch.randn(B, HW, C, device=dev)
    dy = torch.randn(B, HW, C, device=dev)
    w, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
    dx = torch.empty(B, HW, C, device=dev)
    sf = torch.zeros(B * 64, device=dev, dtype=torch.float64)
    sb = torch.zeros(B * 64, device=dev, dtype=torch.float64)
    sp = _lib.stream_ptr()

    def fwd():
        _lib.check(lib.s2i_groupnorm_forward(x.data_ptr(), C, B, HW, C, w.data_ptr(), b.data_ptr(), 1e-5, 1, out.data_ptr(), C,
                                             None, 0, sf.data_ptr(), sp))

    def bwd():
        _lib.check(lib.s2i_groupnorm_backward(dy.data_ptr(), C, x.data_ptr(), C, B, HW, C, w.data_ptr(), b.data_ptr(), 1e-5, 1,
                                              sf.data_ptr(), sb.data_ptr(), None, 0, dx.data_ptr(), C, None, 0, sp))

    res = []
    for fn in (fwd, bwd):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1000 / n)
    return res


def main():
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    if "--sweep" in sys.argv:
        return sweep(lib, dev)
    if "--one" in sys.argv:      # under ncu: a few launches of one shape
        B, HW, C = (int(v) for v in sys.argv[sys.argv.index("--one") + 1:][:3])
        print(time_shape(lib, dev, B, HW, C, 3))
        return
    print("form:", "grid-barrier" if os.environ.get("S2I_GN_CLUSTER") == "0" else "cluster")
    print("%-18s %10s %10s %12s %12s" % ("B x HW x C", "fwd us", "bwd us", "fwd GB/s", "bwd GB/s"))
    for B, HW, C in SHAPES:
        f, b = time_shape(lib, dev, B, HW, C)
        n_el = B * HW * C
        print("%-18s %10.1f %10.1f %12.0f %12.0f" % ("%dx%dx%d" % (B, HW, C), f, b, n_el * 6 / f / 1e3, n_el * 12 / b / 1e3))


def sweep(lib, dev):
    """Every legal (groups per cluster, CTAs per cluster) geometry per shape: fwd/bwd us.  The library reads S2I_GN_GEOM on
    every call."""
    geoms = [(g, c) for g in (1, 2, 4, 8, 16, 32) for c in (1, 2, 4, 8, 16)]
    for B, HW, C in SHAPES:
        os.environ.pop("S2I_GN_GEOM", None)
        f, b = time_shape(lib, dev, B, HW, C, 30)
        print("%dx%dx%d  chosen: %.1f / %.1f" % (B, HW, C, f, b))
        row = []
        for gpc, cs in geoms:
            W = gpc * C // 32
            if W % 4 or W > 1024 or cs > HW:
                continue
            os.environ["S2I_GN_GEOM"] = "%d,%d" % (gpc, cs)
            f, b = time_shape(lib, dev, B, HW, C, 30)
            row.append("(%d,%d) %.1f/%.1f" % (gpc, cs, f, b))
        for i in range(0, len(row), 6):
            print("    " + "   ".join(row[i:i + 6]))
    os.environ.pop("S2I_GN_GEOM", None)


if __name__ == "__main__":
    main()
