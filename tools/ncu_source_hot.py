#!/usr/bin/env python
"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
Prints, per kernel, the source lines ranked by executed warp instructions and by stall samples."""
import csv
import sys
import collections


def main(path, top=40):
    kernels = collections.OrderedDict()
    cur_file, cur_fn, hdr = None, None, None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1]; continue
        if row[0] == "Function Name":
            cur_fn = row[1]; continue
        if row[0] == "Line No":
            hdr = row; continue
        if hdr is None or len(row) < 8:
            continue
        if row[0] == "":         # SASS row
            continue
        try:
            line = int(row[0])
        except ValueError:
            continue
        k = kernels.setdefault(cur_fn, collections.defaultdict(lambda: [0, 0, ""]))
        key = (cur_file.split("/")[-1], line)
        try:
            ni = int(row[hdr.index("Instructions Executed")] or 0); ns = int(row[hdr.index("# Samples")] or 0)
        except ValueError:       # source text with embedded quotes/commas (inline asm in CUDA headers)
            continue
        k[key][0] += ni
        k[key][1] += ns
        k[key][2] = row[1].strip()[:110]
    for fn, lines in kernels.items():
        tot_i = sum(v[0] for v in lines.values()); tot_s = sum(v[1] for v in lines.values())
        print(f"\n=== {fn[:100]}  total warp-inst {tot_i:,}  samples {tot_s:,}")
        for (f, l), (ni, ns, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"{100 * ni / max(tot_i, 1):5.1f}% inst {100 * ns / max(tot_s, 1):5.1f}% stall  {f}:{l:<4d} {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
