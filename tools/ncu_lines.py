#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep (built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_frame_eval [top_n]
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    data, cur_file, hdr = [], "", None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
            ie, ws = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) > ie and r[0].isdigit() and r[ie].isdigit():
            data.append((int(r[ie]), int(r[ws]) if r[ws].isdigit() else 0, cur_file, r[0], r[1]))
    tot_i = sum(d[0] for d in data) or 1
    tot_s = sum(d[1] for d in data) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for d in sorted(data, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100.0 * d[0] / tot_i, 100.0 * d[1] / tot_s,
                                                         d[2], d[3], d[4][:100]))


if __name__ == "__main__":
    main()
