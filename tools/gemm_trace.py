"""GPU box: phase timeline of gemm_tma_kernel for a few shapes: per-CTA medians of the SM cycle counter (clock64) at each phase,
relative to the CTA's own entry (cycles; ~1.9 cycles per ns)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib as L  # noqa: E402
from tools.gemm_bench import SHAPES, make  # noqa: E402

NAMES = ["entry", "setup", "loads_issued", "mmas_issued", "epi_start", "accum_ready", "res_ready", "chunks_done", "stores_read", "exit",
         "load0", "load1", "load2", "chunk0", "chunk1", "chunk2"]
# split-K launches reuse slots: res_ready = partial tile parked, chunks_done = cluster barrier passed, chunk0 = reduction starts,
# chunk1 = first trip's loads issued, chunk2 = first trip finished, stores_read = rows written


def main():
    lib = L.lib()
    lib.s2i_gemm_set_trace.argtypes = [C.c_void_p]
    buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
    for shape in SHAPES:
        want = sys.argv[1:] or ["lin 320->320 @64 res", "conv 320@64", "qkv 320->1152 @64", "conv 1280@16", "conv 1280@8", "lin 1280->1280 @8 res", "lin 1280->1280 @16 res", "conv 640@32"]
        if shape[0] not in want:
            continue
        kw, keep = make(shape)
        d = L.GemmDesc(**kw)
        for _ in range(3):
            L.gemm(d)
        torch.cuda.synchronize()
        buf.zero_()
        lib.s2i_gemm_set_trace(buf.data_ptr())
        L.gemm(d)
        torch.cuda.synchronize()
        lib.s2i_gemm_set_trace(None)
        t = buf.view(-1, 16).cpu()
        t = t[t[:, 0] > 0][:, :16].double()
        t0 = t[:, 0].min()
        rel = t - t0
        rel[t == 0] = float("nan")
        med = rel.nanmedian(dim=0).values
        mx = torch.nan_to_num(rel, nan=0.0).max(dim=0).values
        t = torch.where(t == 0, torch.full_like(t, float("nan")), t)
        print(f"{shape[0]}: {t.shape[0]} CTAs")
        print("   per-CTA median cycles from own entry: " + ", ".join(f"{n}={v:.0f}" for n, v in zip(NAMES, (t - t[:, :1]).nanmedian(dim=0).values.tolist())))


if __name__ == "__main__":
    main()
