"""Condense an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.

usage: python scripts/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr_i]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        agg.setdefault(r[ki], []).append(v * scale)
    total = sum(sum(v) for v in agg.values())
    print(f"# {sys.argv[1]}: ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised)")
    print(f"# {'kernel':78s} launches   mean us   total us  share")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:80]:80s} {len(v):6d} {sum(v) / len(v):9.1f} {sum(v):10.1f} {100 * sum(v) / total:5.1f}%")


if __name__ == "__main__":
    main()
