#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
Usage: python tools/summarize_launches.py launches.csv [out.txt]"""
import collections
import csv
import re
import sys

src = sys.argv[1]
lines = [l for l in open(src) if not l.startswith("==")]
tot, cnt = collections.defaultdict(float), collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
out = [f"# {src}: {sum(cnt.values())} launches, {T / 1e6:.2f} ms of GPU time (ncu: cold caches, serialised -- compare SHARES)"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
    out.append(f"{v / 1e6:10.3f} ms {100 * v / T:5.1f}%  n={cnt[k]:6d}  avg {v / cnt[k] / 1e3:9.1f} us  {k[:100]}")
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
