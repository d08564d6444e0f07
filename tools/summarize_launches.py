"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: python tools/summarize_launches.py file.csv [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0]), 0
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("bg::<unnamed>::", "")[-48:]
    agg[name][0] += 1
    agg[name][1] += v
    n += 1
tot = sum(v[1] for v in agg.values())
print(f"{n} launches, {tot / 1e3:.3f} ms of kernel time (ncu: serialised, cold caches; compare shares)")
print(f"{'kernel':50s} {'n':>5s} {'us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:50s} {v[0]:5d} {v[1]:10.1f} {100 * v[1] / tot:6.1f}%")
