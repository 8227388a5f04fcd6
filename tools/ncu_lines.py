#!/usr/bin/env python3
"""Per-source-line hot spots of one .ncu-rep (needs -lineinfo at compile time and --import-source on at capture):
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [topN]
Aggregates warp-stall samples and executed instructions per CUDA source line (cuda,sass correlated view)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = defaultdict(lambda: [0, 0, ""])
    cur_file, hdr = None, None
    first_kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            if first_kernel is None:
                first_kernel = r[1]
            active = (r[1] == first_kernel)
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not active:
            continue
        if r[0] not in ("", "-") and r[0].isdigit():
            d = dict(zip(hdr, r))
            key = (cur_file, int(r[0]))
            try:
                agg[key][0] += int(r[4]); agg[key][1] += int(r[7])
            except ValueError:
                pass
            agg[key][2] = r[1].strip()[:110]
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    print(first_kernel)
    print(f"total stall samples {tot_s}, total warp instructions {tot_i}")
    for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% inst  {f}:{ln:<4d} {src}")


if __name__ == "__main__":
    main()
