#!/usr/bin/env python
"""Instruction mix of one kernel from `cuobjdump -sass` (static counts; the frame loop of the
gram kernels is straight-line, so static counts ~ per-frame counts plus prologue).
usage: tools/sass_stats.py <object-or-so> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            counts[m.group(1)] += 1
    tot = sum(counts.values())
    print(f"{pat}: {tot} instructions")
    for k, v in counts.most_common(40):
        print(f"  {k:12s} {v}")


if __name__ == "__main__":
    main()
