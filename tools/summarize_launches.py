"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=45):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r"<unnamed>::", "", name)[:100]
        tot[name][0] += 1
        tot[name][1] += v
        n += 1
    T = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if "destr::" in k)
    print(f"launches {n}  total {T:.1f} us  (destr:: kernels {ours:.1f} us = {100*ours/T:.1f}%)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:9.1f} us {100*v[1]/T:5.1f}%  x{v[0]:4d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
