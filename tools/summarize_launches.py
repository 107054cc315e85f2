"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (gpurun_out/<tag>/launches.csv) into the
text committed under profiles/: per-kernel totals and the launch sequence of the last (unguided) denoising step.
usage: python tools/summarize_launches.py gpurun_out/r1a/launches.csv > profiles/r1_launches_v1.txt"""
import collections
import csv
import re
import sys


def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
    seq = []
    for r in rows:
        m = re.search(r"(\w+)\s*(<[^(]*)?\(", r[ki])
        seq.append((m.group(1) if m else r[ki][:40], r[gi], float(r[vi]) / 1e3))
    return seq


def main(path):
    seq = load(path)
    tot = sum(s[2] for s in seq)
    print(f"# {path}: {len(seq)} launches, {tot / 1e3:.3f} ms (ncu per-launch times: cold cache, serialised -- compare shares)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, us in seq:
        agg[n][0] += 1
        agg[n][1] += us
    print(f"{'kernel':34s} {'launches':>8s} {'ms':>9s} {'avg_us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:34s} {v[0]:8d} {v[1] / 1e3:9.3f} {v[1] / v[0]:8.1f} {v[1] / tot:6.3f}")
    starts = [i for i, s in enumerate(seq) if s[0] == "timestep_embedding_kernel"]
    if starts:
        s0 = starts[-1]
        print(f"\n# launch sequence of the last denoising step ({len(seq) - s0} launches, {sum(s[2] for s in seq[s0:]) / 1e3:.3f} ms)")
        for i in range(s0, len(seq)):
            print(f"{i - s0:4d} {seq[i][0]:30s} {seq[i][1]:16s} {seq[i][2]:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
