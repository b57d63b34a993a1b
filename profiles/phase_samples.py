#!/usr/bin/env python3
"""Split the warp-stall samples of an `ncu --page source --csv` dump of one of the MLP tile kernels into its phases:
the DMMA loops (GEMMs) and the code between them.  usage: phase_samples.py source.csv"""
import collections
import csv
import sys

csv.field_size_limit(1 << 30)


def main(path):
    rows = list(csv.reader(open(path)))[2:]
    tot = sum(int(r[2]) for r in rows)
    idx = [i for i, r in enumerate(rows) if "DMMA" in r[1]]
    groups, start, prev = [], idx[0], idx[0]
    for i in idx[1:]:
        if i - prev > 60:
            groups.append((start, prev))
            start = i
        prev = i
    groups.append((start, prev))
    bounds = [0]
    for a, b in groups:
        bounds += [a, b + 1]
    bounds.append(len(rows))
    gemm = 0.0
    for k in range(len(bounds) - 1):
        a, b = bounds[k], bounds[k + 1]
        sm = sum(int(r[2]) for r in rows[a:b])
        ex = sum(int(r[5]) for r in rows[a:b])
        ops = collections.Counter()
        for r in rows[a:b]:
            s = r[1].split()
            o = s[1] if s[0].startswith("@") else s[0]
            ops[o.split(".")[0]] += int(r[2])
        kind = "GEMM " if k % 2 == 1 else "phase"
        gemm += sm if k % 2 == 1 else 0
        print(f"{kind} sass {a:5d}-{b:5d}  samples {100 * sm / tot:5.1f} %  warp-instructions {ex:11d}  "
              f"top {[(o, round(100 * v / tot, 1)) for o, v in ops.most_common(4)]}")
    print(f"in DMMA loops: {100 * gemm / tot:.1f} % of the samples")


if __name__ == "__main__":
    main(sys.argv[1])
