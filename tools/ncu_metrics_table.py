#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics ... --csv --log-file X.csv` launch list: mean of every metric per kernel name,
launch count and totals for the first launch set (one step).  Usage: python tools/ncu_metrics_table.py X.csv [first_n_launches_per_kernel]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    per = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    hdr = None
    data = collections.OrderedDict()
    seen = collections.Counter()
    ids = {}
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(LfDev")[0].replace("void ", "")
        key = (name, d["ID"])
        if key not in ids:
            seen[name] += 1
            ids[key] = seen[name]
        if ids[key] > per:
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        if d["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else v if u == "us" else v * 1e3 if u == "ms" else v
        data.setdefault(name, collections.defaultdict(float))[d["Metric Name"]] += v
    names = sorted({m for v in data.values() for m in v})
    print("kernel".ljust(34) + "".join(n.split("__")[-1][:26].rjust(28) for n in names))
    tot = collections.defaultdict(float)
    for k, v in data.items():
        print(k[:33].ljust(34) + "".join(("%.4g" % v[n]).rjust(28) for n in names))
        for n in names:
            tot[n] += v[n]
    print("TOTAL".ljust(34) + "".join(("%.4g" % tot[n]).rjust(28) for n in names))


if __name__ == "__main__":
    main()
