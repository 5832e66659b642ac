#!/usr/bin/env python
"""Per-source-line warp-stall samples of a kernel from an .ncu-rep captured with --import-source on (and -lineinfo).

    python tools/ncu_hot_lines.py gpurun_out/r01b/prof_ce.ncu-rep [top N] > profiles/<name>_hot_lines.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    file_name, hdr = None, None
    per_line = defaultdict(lambda: [0.0, 0.0, 0.0, ""])  # samples, instructions, thread instructions, text
    per_file = defaultdict(float)
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            file_name = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_samples, i_inst, i_thr = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or len(r) <= i_thr or r[2] != "-":  # keep the per-line aggregate rows (Address == "-")
            continue
        try:
            s, n, t = float(r[i_samples]), float(r[i_inst]), float(r[i_thr])
        except ValueError:
            continue
        key = (file_name, int(r[0]))
        per_line[key][0] += s
        per_line[key][1] += n
        per_line[key][2] += t
        per_line[key][3] = r[1].strip()
        per_file[file_name] += s
    total = sum(v[0] for v in per_line.values()) or 1.0
    print(f"total warp-stall samples: {total:.0f}")
    print("by file:")
    for f, s in sorted(per_file.items(), key=lambda kv: -kv[1]):
        print(f"  {100 * s / total:5.1f}%  {f}")
    print(f"top {top} source lines (samples %, warp instructions, avg active threads):")
    for (f, line), (s, n, t, text) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100 * s / total:5.1f}%  {n:14.0f}  {t / n if n else 0:5.1f}  {f}:{line}  {text[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
