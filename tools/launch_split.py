#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): count, total ms, first/last us.

    python tools/launch_split.py gpurun_out/ev4/launches_event.csv
"""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, seq = defaultdict(float), defaultdict(list)
    for r in rows[start:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].split("(")[0][-34:]
        tot[name] += v
        seq[name].append(v)
    total = sum(tot.values())
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        s = seq[k]
        print(f"{k:36s} n={len(s):5d} total={v / 1e6:9.3f} ms ({100 * v / total:5.1f} %)  first us {[round(x / 1e3) for x in s[:8]]} last {[round(x / 1e3) for x in s[-4:]]}")


if __name__ == "__main__":
    main(sys.argv[1])
