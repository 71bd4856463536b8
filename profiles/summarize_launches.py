"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_bench_launches.txt
"""
import collections
import csv
import re
import sys


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    i_name, i_metric, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= i_val or r[i_metric] != "gpu__time_duration.sum":
            continue
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        if v != v:
            continue
        unit = r[i_unit]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((r[i_name], us))
    tot = sum(u for _, u in rows)
    agg = collections.OrderedDict()
    for name, us in rows:
        key = re.sub(r"\(.*", "", name).replace("void ", "")[:90]
        n, s = agg.get(key, (0, 0.0))
        agg[key] = (n + 1, s + us)
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time (ncu per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"{'kernel':92s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
    for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{key:92s} {n:8d} {s:10.1f} {s / n:8.2f} {100 * s / tot:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
