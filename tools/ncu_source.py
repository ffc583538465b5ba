#!/usr/bin/env python
"""Rank the CUDA source lines of one kernel in an ncu report by warp instructions executed.

usage: tools/ncu_source.py REPORT.ncu-rep KERNEL_REGEX[:launch-index] [top]
Needs a capture made with --import-source on of a build with -lineinfo.
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    name, _, inst = kern.partition(":")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-id", "::regex:^%s$:%s" % (name, inst or "1")], capture_output=True, text=True).stdout
    rows, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "File Path":
            cur = row[1].split("/")[-1]
            continue
        if len(row) > 8 and row[0].isdigit() and row[7] not in ("-", ""):
            rows.append((cur, int(row[0]), row[1].strip()[:100], int(row[6] or 0), int(row[7]), int(row[8])))
    ti = sum(r[4] for r in rows) or 1
    tt = sum(r[5] for r in rows)
    ts = sum(r[3] for r in rows) or 1
    print("warp-instr %d  thread-instr %d  threads/instr %.1f  samples %d" % (ti, tt, tt / ti, ts))
    print("%-16s %5s %7s %7s %5s  %s" % ("file", "line", "instr%", "stall%", "thr", "source"))
    for f, l, s, sm, i, t in sorted(rows, key=lambda r: -r[4])[:top]:
        print("%-16s %5d %7.2f %7.2f %5.1f  %s" % (f[:16], l, 100.0 * i / ti, 100.0 * sm / ts, t / max(i, 1), s))


if __name__ == "__main__":
    main()
