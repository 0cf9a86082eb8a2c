#!/usr/bin/env python
"""Kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list (read here, no GPU needed).
usage: python tools/launch_shares.py gpurun_out/x_launches.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    c, t = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= iV:
            continue
        v = float(r[iV].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1e-3)
        k = r[iK].split("(")[0].replace("b2::", "")
        if "at::" in k or "vectorized" in k or "elementwise" in k:
            k = "(torch) " + k.split("<")[0].split("::")[-1]
        c[k] += 1; t[k] += v
    tot = sum(t.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total us':>11s} {'avg us':>9s} {'share':>7s}")
    for k, v in t.most_common():
        print(f"{k[:44]:44s} {c[k]:8d} {v:11.1f} {v / c[k]:9.2f} {100 * v / tot:6.1f}%")
    print(f"{'all':44s} {sum(c.values()):8d} {tot:11.1f}")


if __name__ == "__main__":
    main()
