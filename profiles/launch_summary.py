#!/usr/bin/env python
"""Summarise ncu launch lists (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`):
one line per launch -- kernel, grid, duration, DRAM bytes read / written.

  python profiles/launch_summary.py [--max N] a.csv b.csv ... > profiles/rNN_xxx_launches.txt
"""
import argparse
import csv
import os
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+")
    ap.add_argument("--max", type=int, default=14)
    a = ap.parse_args()
    for f in a.files:
        launches = {}
        for r in csv.reader(open(f, errors="replace")):
            if len(r) < 15 or not r[0].isdigit():
                continue
            d = launches.setdefault(int(r[0]), {"kernel": r[4], "grid": r[8]})
            d[r[12]] = float(r[14].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3,
                                                          "Mbyte": 1e6, "Gbyte": 1e9}.get(r[13], 1.0)
        print("== %s (%d launches)" % (os.path.basename(f), len(launches)))
        tot = 0.0
        for i in sorted(launches)[:a.max]:
            d = launches[i]
            name = re.sub(r"\(.*", "", d["kernel"].replace("void ", ""))
            us = d.get("gpu__time_duration.sum", 0.0)
            tot += us
            print("%-3d %-42s %-14s %8.1f us  dram rd %8.1f MB wr %8.1f MB" % (
                i, name[:42], d["grid"], us, d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
        print("sum of the listed launches: %.1f us" % tot)


if __name__ == "__main__":
    main()
