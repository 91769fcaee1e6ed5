"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share).

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot:.3f} ms of kernel time (cold-cache, serialised under ncu: compare SHARES)")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:64]:64s} n={c:5d} total={ms:9.3f} ms avg={ms / c * 1e3:9.1f} us share={ms / tot:6.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
