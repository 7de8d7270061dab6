#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import _path  # noqa: F401  (repository root on sys.path)
import collections
import csv
import io
import sys


def main(path, frames=64):
    text = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r["Kernel Name"].replace("void ", "").replace("unnamed>::", "").split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised; compare SHARES)")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")
    print(f"# per frame ({frames} frames/step): {tot / frames:.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 64)
