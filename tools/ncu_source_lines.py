#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report captured
with --import-source on:  python tools/ncu_source_lines.py report.ncu-rep kernel_regex [top] [launch_skip]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    skip = sys.argv[4] if len(sys.argv) > 4 else "0"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    cur = None
    agg = {}
    fpath = ""
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_ins = hdr.index("Instructions Executed")
            i_smp = hdr.index("# Samples")
            continue
        if hdr is None:
            continue
        if r[0] != "":
            if not r[0].isdigit():
                continue
            cur = (fpath, int(r[0]), r[1].strip())
            agg.setdefault(cur, [0, 0])
            continue
        if cur is None or len(r) <= i_ins:
            continue
        try:
            agg[cur][0] += int(r[i_ins])
            agg[cur][1] += int(r[i_smp])
        except ValueError:
            pass
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    for (f, ln, src), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f}:{ln:5d}  inst {100 * ins / ti:5.1f}%  samples {100 * smp / ts:5.1f}%  {src[:100]}")


if __name__ == "__main__":
    main()
