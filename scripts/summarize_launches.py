"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total time)."""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"pycmf::(?:\(anonymous namespace\)::)?(\w+)(<[^(]*>)?", name)
    if m:
        return "pycmf::" + m.group(1) + (m.group(2) or "")
    return re.sub(r"\(.*", "", name)[:70]


def main(path, skip=0):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot = collections.OrderedDict()
    for i, row in enumerate(csv.DictReader(lines)):
        if i < skip:
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row.get("Metric Unit", "us"), 1.0)
        t = tot.setdefault(short(row["Kernel Name"]), [0, 0.0])
        t[0] += 1
        t[1] += v
    s = sum(v[1] for v in tot.values())
    print("%-78s %6s %12s %7s" % ("kernel", "n", "total us", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-78s %6d %12.1f %6.1f%%" % (k[:78], v[0], v[1], 100 * v[1] / s))
    print("total %.1f us over %d launches" % (s, sum(v[0] for v in tot.values())))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
