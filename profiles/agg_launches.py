"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python agg_launches.py file.csv"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", row["Kernel Name"]))
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"total {tot:.3f} ms over {sum(n for n, _ in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.4f} ms  {k[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])
