"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST step."""
import csv
import re
import sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((name, us))
    return rows


def main():
    path = sys.argv[1]
    first = sys.argv[2] if len(sys.argv) > 2 else None  # kernel-name substring that starts a step
    rows = load(path)
    if first:
        starts = [i for i, (n, _) in enumerate(rows) if first in n]
        rows = rows[starts[-1]:] if starts else rows
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("%-70s %5s %10s %6s" % ("kernel", "n", "us", "share"))
    for n, (c, us) in agg.items():
        print("%-70s %5d %10.1f %5.1f%%" % (n[:70], c, us, 100 * us / tot))
    print("%-70s %5s %10.1f" % ("TOTAL", "", tot))


if __name__ == "__main__":
    main()
