"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).

    python tools/summarize_launches.py gpurun_out/launches.csv "header line" > profiles/rNN_launches_*.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)                      # drop the argument list
    name = name.replace("scrc::", "").replace("void ", "")
    return name[:110]


def main():
    path = sys.argv[1]
    header = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v * 1e6 if unit in ("s", "second") else v
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
        total += us; n += 1
    print(header)
    print("(cold-cache, serialised launches: compare SHARES, not absolutes)")
    print(f"{n} launches, {total / 1e3:.1f} ms of kernel time\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.2f}%  {t:12.1f} us  n={c:6d}  avg {t / c:10.2f} us  {k}")


if __name__ == "__main__":
    main()
