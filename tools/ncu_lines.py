#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per source line:
stall samples and executed instructions, top-N lines.  Usage: ncu_lines.py dump.csv [topN]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cur_file, per = None, {}
    hdr = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1][:60]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr and r[0].isdigit():
            try:
                s, n = int(r[i_s]), int(r[i_i])
            except ValueError:
                continue
            key = (cur_file, int(r[0]))
            a = per.setdefault(key, [0, 0, r[1][:110]])
            a[0] += s
            a[1] += n
    tot_s = sum(v[0] for v in per.values()) or 1
    tot_i = sum(v[1] for v in per.values()) or 1
    print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
    print("%-22s %6s %7s %7s  %s" % ("file:line", "smp%", "inst%", "", "source"))
    for (f, ln), (s, n, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %6.2f %7.2f          %s" % ("%s:%d" % (f, ln), 100.0 * s / tot_s, 100.0 * n / tot_i, src))


if __name__ == "__main__":
    main()
