#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, title=""):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3, "nsecond": 1e-6}.get(unit, 1e-6)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"<unnamed>::", "", name).replace("(anonymous namespace)::", "")
        rows.append((name, ms))
    agg = OrderedDict()
    for n, ms in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + ms)
    total = sum(t for _, t in agg.values())
    if title:
        print(title)
    print("per-launch times are cold-cache and serialised (compare SHARES; concurrent SMO launches overlap in the real step): "
          "launches=%d total=%.1f ms" % (len(rows), total))
    print("%-72s %6s %10s %7s" % ("kernel", "count", "total ms", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %6d %10.2f %6.1f%%" % (n[:72], c, t, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
