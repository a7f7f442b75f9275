#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python tools/summarize_launches.py list.csv [top]"""
import collections
import csv
import re
import sys


def summarize(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("dv3::<unnamed>::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    out = [f"# {n} launches, {tot / 1e3:.2f} ms summed kernel time (cold-cache, serialised: compare SHARES, not absolutes)"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"{t:9.1f} us {100 * t / tot:5.1f}%  n={c:4d} avg={t / c:8.2f} us  {k}")
    return out


if __name__ == "__main__":
    print("\n".join(summarize(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)))
