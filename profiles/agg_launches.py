"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        name = row["Kernel Name"]
        m = re.match(r"(?:void )?(?:vb::)?(\w+)(?:<([^>]*)>)?", name)
        short = m.group(1) + ("<" + m.group(2) + ">" if m and m.group(2) else "") if m else name[:60]
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d} avg={v[1] / v[0]:9.1f} us  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
