"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (last find only)."""
import collections
import csv
import sys


def main(path, marker="k_extract", nth_last=1):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [x["Kernel Name"] for x in rows]
    idx = [i for i, n in enumerate(names) if marker in n]
    start = idx[-int(nth_last)] if len(idx) >= int(nth_last) else 0
    agg = collections.OrderedDict()
    for x in rows[start:]:
        n = x["Kernel Name"].split("(")[0]
        v = float(x["Metric Value"])
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':40s} {'n':>4s} {'us':>10s} {'share':>6s}")
    for n, (c, t) in agg.items():
        print(f"{n:40s} {c:4d} {t:10.1f} {100 * t / tot:5.1f}%")
    print(f"{'total':40s} {sum(a[0] for a in agg.values()):4d} {tot:10.1f}")


if __name__ == "__main__":
    main(*sys.argv[1:])
