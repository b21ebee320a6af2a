#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum ... --csv) into a markdown table.

    python tools/launch_list.py gpurun_out/launches_c2.csv "title" > profiles/rNN_launches_<cfg>.md
"""
import collections
import csv
import re
import sys


def main():
    path, title = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*$", "", r[ik])
        name = name if len(name) < 72 else name[:72]
        us = float(r[iv].replace(",", "")) / (1e3 if r[iu] == "ns" else 1.0)
        acc.setdefault(name, []).append(us)
    print(f"# {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv` under gpurun, one "
          "B200. Per-launch times are cold-cache and serialised: compare shares, not absolutes.\n")
    print("| kernel | launches | mean us | min us | max us | total us |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in acc.items():
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {max(v):.2f} | {sum(v):.1f} |")


if __name__ == "__main__":
    main()
