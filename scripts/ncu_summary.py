"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) count/mean/total."""
import collections
import csv
import sys

path = sys.argv[1]
after = sys.argv[2] if len(sys.argv) > 2 else None      # only the launches from the first kernel whose name contains this
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(list)
started = after is None
for row in csv.DictReader(lines):
    started = started or after in row["Kernel Name"]
    if not started:
        continue
    agg[(row["Kernel Name"][:64], row["Grid Size"])].append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':66s}{'grid':18s}{'n':>5s}{'mean ns':>12s}{'total us':>12s}{'share':>8s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[0]:66s}{k[1]:18s}{len(v):5d}{sum(v)/len(v):12.1f}{sum(v)/1e3:12.1f}{100*sum(v)/tot:7.1f}%")
