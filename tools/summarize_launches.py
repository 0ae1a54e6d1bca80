"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, skip_build=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        agg[row["Kernel Name"]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%10s %5s %9s %6s  kernel" % ("total_us", "n", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%10.1f %5d %9.1f %5.1f%%  %s" % (sum(v), len(v), sum(v) / len(v), 100 * sum(v) / tot, k[:110]))


if __name__ == "__main__":
    main(sys.argv[1])
