"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
Usage: python scripts/summarize_launches.py gpurun_out/launches_train.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"):
            v *= 1e3
        elif unit in ("ms", "msecond"):
            v *= 1e6
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        rows.append((int(r["ID"]), name, v))
    rows = rows[skip:]
    agg = OrderedDict()
    for _, n, v in rows:
        c, t = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, t + v)
    total = sum(t for _, t in agg.values())
    print(f"{len(rows)} launches, total {total / 1e6:.3f} ms")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:10.1f} us  {100 * t / total:5.1f}%  x{c:<4d} avg {t / c / 1e3:8.1f} us  {n}")


if __name__ == "__main__":
    main()
