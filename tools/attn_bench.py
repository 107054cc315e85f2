"""GPU box: device time of the fused attention forward / backward kernels on the SD1.5 CFG-pair shapes (graph-timed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib as L  # noqa: E402

SHAPES = [("self N=4096 d=40", 2, 8, 4096, 4096, 40, 48), ("self N=1024 d=80", 2, 8, 1024, 1024, 80, 80),
          ("self N=256 d=160", 2, 8, 256, 256, 160, 160), ("cross N=1024 Nk=77 d=80", 2, 8, 1024, 77, 80, 80),
          ("self N=4096 d=40 B=1", 1, 8, 4096, 4096, 40, 48),
          ("cross N=4096 Nk=77 d=40", 2, 8, 4096, 77, 40, 48), ("self N=9216 d=64 (SD2.1)", 2, 5, 9216, 9216, 64, 64)]


def graph_time(fn, reps=10):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def sweep(lib):
    """Forward only: shared-memory ring / score-tile configurations per shape."""
    for name, B, heads, Nq, Nk, d, dp in SHAPES:
        HP = heads * dp
        gen = torch.Generator().manual_seed(1)
        q = torch.randn(B, Nq, HP, generator=gen).cuda().half()
        kv = torch.randn(B, Nk, 2 * HP, generator=gen).cuda().half()
        out = torch.zeros(B, Nq, HP, device="cuda", dtype=torch.float16)
        lse = torch.zeros(B * heads, Nq, device="cuda")

        def fwd():
            L.check(lib.s2i_attention(q.data_ptr(), HP, 0, kv.data_ptr(), 2 * HP, 0, HP, B, heads, Nq, Nk, dp, d, d ** -0.5,
                                      out.data_ptr(), HP, lse.data_ptr(), L.stream_ptr()))

        row = []
        for cfg in ("", "2,3,2", "2,3,1", "2,2,1", "1,4,2", "1,4,1", "1,3,1", "1,2,1"):
            if cfg:
                os.environ["S2I_ATTN_CFG"] = cfg
            else:
                os.environ.pop("S2I_ATTN_CFG", None)
            row.append("%s %.1f" % (cfg or "default", graph_time(fwd)))
        print("%-28s (pbufs,stages,nS) us: %s" % (name, "   ".join(row)), flush=True)
    os.environ.pop("S2I_ATTN_CFG", None)


def main():
    lib = L.lib()
    if "--sweep" in sys.argv:
        return sweep(lib)
    for name, B, heads, Nq, Nk, d, dp in SHAPES:
        HP = heads * dp
        gen = torch.Generator().manual_seed(1)
        q = torch.randn(B, Nq, HP, generator=gen).cuda().half()
        kv = torch.randn(B, Nk, 2 * HP, generator=gen).cuda().half()
        dO = torch.randn(B, Nq, HP, generator=gen).cuda().half()
        out = torch.zeros(B, Nq, HP, device="cuda", dtype=torch.float16)
        lse = torch.zeros(B * heads, Nq, device="cuda")
        delta = torch.zeros(B * heads, Nq, device="cuda")
        dq = torch.zeros(B, Nq, HP, device="cuda", dtype=torch.float16)
        dkv = torch.zeros(B, Nk, 2 * HP, device="cuda", dtype=torch.float16)
        scale = d ** -0.5

        def fwd():
            L.check(lib.s2i_attention(q.data_ptr(), HP, 0, kv.data_ptr(), 2 * HP, 0, HP, B, heads, Nq, Nk, dp, d, scale,
                                      out.data_ptr(), HP, lse.data_ptr(), L.stream_ptr()))

        def bwd(with_kv):
            L.check(lib.s2i_attention_backward(q.data_ptr(), HP, 0, kv.data_ptr(), 2 * HP, 0, HP, dO.data_ptr(), out.data_ptr(), HP,
                                               lse.data_ptr(), delta.data_ptr(), B, heads, Nq, Nk, dp, d, scale, dq.data_ptr(), HP, 0,
                                               dkv.data_ptr() if with_kv else None, 2 * HP, 0, HP, L.stream_ptr()))

        tf = graph_time(fwd)
        bwd_ok = dp <= 192
        tq = graph_time(lambda: bwd(False)) if bwd_ok else float("nan")
        tkv = graph_time(lambda: bwd(True)) if Nk >= 128 and bwd_ok else float("nan")
        gf = 4.0 * B * heads * Nq * Nk * d / 1e9
        print(f"{name:28s} fwd {tf:7.1f} us ({gf / tf:6.1f} TF/s)   bwd dQ {tq:7.1f} us   bwd dQ+dK+dV {tkv:7.1f} us", flush=True)


if __name__ == "__main__":
    main()
