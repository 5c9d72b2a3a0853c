"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total device time, share.
usage: launch_shares.py launches.csv [> profiles/rNN_launch_shares_<model>.txt]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("b200::", "")
    name = re.sub(r"at::native::.*", "ATen elementwise / fill / copy", name)
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {len(rows)} launches, {tot:.1f} ms of device time (ncu: serialised, cold-cache: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:92s} n={v[0]:5d} {v[1]:9.2f} ms {100 * v[1] / tot:5.1f}%")
