"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total us and share."""
import collections
import csv
import sys


def main(path, title=""):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    if title:
        print("# " + title)
    print("# times are cold-cache/serialised under ncu: compare SHARES, not absolutes")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s launches=%4d  %10.1f us  %5.1f%%" % (k[:44], n, t, 100 * t / tot))
    print("total us %.1f over %d launches" % (tot, sum(a[0] for a in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
