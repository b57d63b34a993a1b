#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list: per-kernel launch count, total time, share, DRAM bytes.  usage: summarize_launches.py list.csv [header text]"""
import collections
import csv
import re
import sys


def main(path, header=""):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    t = collections.Counter()
    n = collections.Counter()
    rd = collections.Counter()
    wr = collections.Counter()
    ids = collections.defaultdict(set)
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void ", "", name)
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit, metric = r[ix["Metric Unit"]], r[ix["Metric Name"]]
        if metric == "gpu__time_duration.sum":
            v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
            t[name] += v
            ids[name].add(r[ix["ID"]])
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            (rd if "read" in metric else wr)[name] += v
    for k in ids:
        n[k] = len(ids[k])
    T = sum(t.values())
    if header:
        print(header)
    print(f"total ms {T / 1e6:.1f}  launches {sum(n.values())}")
    for k, v in t.most_common():
        if v / T < 0.0005:
            continue
        line = f"{k[-62:]:62s} n={n[k]:5d} ms={v / 1e6:9.2f} share={100 * v / T:5.1f}%"
        if rd or wr:
            line += f"  dram read {rd[k] / 1e9:7.2f} GB  write {wr[k] / 1e9:7.2f} GB"
        print(line)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
