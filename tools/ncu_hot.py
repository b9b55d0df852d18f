#!/usr/bin/env python
"""Hottest SASS instructions (stall samples) of one kernel in an .ncu-rep captured with --import-source on.
   python tools/ncu_hot.py file.ncu-rep kernel_regex [launch_index] [top_n]"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(io.StringIO(out))]
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hi]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for n, r in enumerate(rows[hi + 1:]):
        try:
            data.append((int(r[i_s] or 0), int(r[i_ex] or 0), n, r[i_src].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    totex = sum(d[1] for d in data)
    print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "", "| samples", tot, "warp-instr", totex, "sass lines", len(data))
    for s, ex, n, src in sorted(data, key=lambda x: -x[0])[:top]:
        print("%5d %7d %5.1f%% ex %11d  %s" % (n, s, 100.0 * s / tot, ex, src[:110]))


if __name__ == "__main__":
    main()
