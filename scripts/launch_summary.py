"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:64]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:64s} n={n:4d} total={t:9.1f}us avg={t / n:7.1f}us share={t / tot:6.1%}")
