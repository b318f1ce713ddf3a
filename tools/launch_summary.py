"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        agg.setdefault(row["Kernel Name"].split("(")[0][-48:], []).append(ms)
    tot = sum(sum(v) for v in agg.values())
    print("kernel,launches,total_ms,share,per_launch_ms")
    for k, v in agg.items():
        print(f"{k},{len(v)},{sum(v):.3f},{sum(v) / tot:.3f},{sum(v) / len(v):.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
