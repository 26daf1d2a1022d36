"""Aggregates an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel: time, share, launches.
usage: python tools/summarize_launches.py gpurun_out/l_resnet18.csv [--all]"""
import collections
import csv
import re
import sys


def load(path):
    hdr, seq = None, []
    for r in csv.reader(open(path)):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        seq.append((re.sub(r"\(.*", "", d["Kernel Name"]), float(d["Metric Value"].replace(",", "")), d.get("Grid Size", "")))
    return seq


def main():
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t, _ in seq:
        agg[n][0] += 1
        agg[n][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"{sys.argv[1]}: {len(seq)} launches, {tot / 1e6:.3f} ms (cold-cache, serialised under ncu)")
    print("| ms | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% | {v[0]} | `{k}` |")
    if "--all" in sys.argv:
        for i, (n, t, g) in enumerate(seq):
            print(i, f"{t / 1e3:9.1f} us", n, g)


if __name__ == "__main__":
    main()
