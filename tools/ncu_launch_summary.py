#!/usr/bin/env python
"""Aggregate an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file LIST.csv ...) per kernel:
   python tools/ncu_launch_summary.py LIST.csv OUT.csv ["comment line"]"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").strip()
        name = re.sub(r"^void ", "", name)
        if name.startswith("cub::"):
            name = re.sub(r"<.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        if note:
            f.write("# " + note + "\n")
        f.write("kernel,launches,total_ms,share_pct,avg_ms_per_launch\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.3f,%.1f,%.4f\n' % (k, n, ms, 100.0 * ms / total, ms / n))
    print("wrote", dst, "total %.3f ms over %d launches" % (total, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main()
