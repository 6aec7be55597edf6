"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv"""
import collections
import csv
import re
import sys


def summarise(path, top=40):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1e-3)
        m = re.search(r"(\w+_kernel|vectorized_elementwise_kernel|\w+Kernel\w*)", name)
        key = m.group(1) if m else name[:60]
        t = re.search(r"_kernel<([^>]*)>", name)
        if t:
            key += "<" + t.group(1)[:30] + ">"
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    out = [f"{path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised)"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        out.append(f"  {t:10.0f} us {100 * t / tot:5.1f}%  n={n:4d} avg={t / n:8.1f} us  {k[:90]}")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarise(p))
