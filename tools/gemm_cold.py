"""GPU box: does it matter that the weights of a launch come from HBM?  Times the K-heavy / weight-streaming GEMM shapes with
ONE weight tensor reused by every launch (L2-warm, what tools/gemm_bench.py measures) and with the launches cycling through
enough distinct weight tensors to exceed the 126 MB L2 (HBM-cold weights, warm activations: the situation inside a step)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib as L  # noqa: E402
from tools.gemm_bench import SHAPES, make  # noqa: E402


def time_many(descs, reps=24):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for d in descs:
            L.gemm(d)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for i in range(reps):
                L.gemm(descs[i % len(descs)])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def main():
    print(f"{'shape':28s} {'weights MB':>10s} {'warm us':>8s} {'cold us':>8s}")
    for shape in SHAPES:
        name, side, rows, N, K, taps = shape[:6]
        wbytes = N * taps * K * 2
        if wbytes < 3e6:
            continue
        kw, keep = make(shape)
        ncopies = max(2, int(300e6 // wbytes) + 1)
        ws = [keep[1]] + [keep[1].clone() for _ in range(ncopies - 1)]
        descs = [L.GemmDesc(**{**kw, "B": w.data_ptr()}) for w in ws]
        warm = time_many(descs[:1])
        cold = time_many(descs)
        print(f"{name:28s} {wbytes / 1e6:10.1f} {warm:8.1f} {cold:8.1f}", flush=True)
        del keep, ws


if __name__ == "__main__":
    main()
