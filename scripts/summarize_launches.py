"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (profiles/*.csv[.gz])."""
import collections
import csv
import gzip
import re
import sys


def main(path, skip_frac=0.0):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", ""), v))
    rows = rows[int(len(rows) * skip_frac):]
    tot = sum(v for _, v in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% | %.2f |" % (k[:70], c, t, 100 * t / tot, t / c))
    print("\ntotal: %d launches, %.1f us (cold-cache, serialised by ncu: compare shares, not absolutes)" % (len(rows), tot))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
