#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_<tag>.md
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("kzg::", "").replace("void ", "")
    return name


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        ns = float(r[14].replace(",", ""))
        key = (short(r[4]), r[8], r[7])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    print("launches: %d, total device time %.1f ms (ncu: serialised, cold cache -- compare shares)\n" % (len(rows), total / 1e6))
    print("| kernel | grid | block | launches | total ms | share | avg ms |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (name, grid, block), (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %s | %s | %d | %.2f | %.1f%% | %.3f |" % (name, grid, block, cnt, ns / 1e6, 100 * ns / total, ns / 1e6 / cnt))


if __name__ == "__main__":
    main()
