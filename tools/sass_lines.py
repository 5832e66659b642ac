#!/usr/bin/env python
"""Static SASS instruction count per source line / per inlined function of one kernel (needs -lineinfo).

    python tools/sass_lines.py minimc_b200/csrc/build/kernels.o 'fixed_source_kernelILi0ELb1ELb0' [top N]

Answers "where do the kernel's N thousand instructions come from" before spending GPU time.
"""
import re
import subprocess
import sys
import tempfile
from collections import Counter
from pathlib import Path


def main(obj, pattern, top=40):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=d, check=True, capture_output=True)
        cubin = next(Path(d).glob("*.cubin"))
        text = subprocess.run(["nvdisasm", "--print-line-info-inline", str(cubin)], capture_output=True, text=True).stdout
    in_kernel = False
    per_line, per_file, inclusive, total = Counter(), Counter(), Counter(), 0
    frames, in_group = [], False  # frames[0] = innermost line of the current instruction group
    for ln in text.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            in_kernel = pattern in ln
            continue
        if not in_kernel:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            if not in_group:
                frames, in_group = [], True
            frames.append((Path(m.group(1)).name, int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
            in_group = False
            total += 1
            if frames:
                per_line[frames[0]] += 1
                per_file[frames[0][0]] += 1
                for f in set(frames):
                    inclusive[f] += 1
    print(f"{total} SASS instructions in kernels matching {pattern!r}")
    for f, n in per_file.most_common():
        print(f"  {n:7d}  {f}")
    print(f"top {top} lines:")
    for (f, l), n in per_line.most_common(top):
        print(f"  {n:6d}  {f}:{l}")
    print(f"top {top} lines, inclusive of everything inlined under them:")
    for (f, l), n in inclusive.most_common(top):
        print(f"  {n:6d}  {f}:{l}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
