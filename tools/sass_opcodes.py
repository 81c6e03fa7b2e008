#!/usr/bin/env python
"""Opcode counts per kernel from `cuobjdump -sass` (which data-movement / synchronisation instructions a kernel really contains).

    python tools/sass_opcodes.py blueberry_b200/lib/hist.o hist_pairs_kernelILb0ELb1ELb1   [object] [substring of the mangled name]
"""
import collections
import re
import subprocess
import sys


def main(obj, needle):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or needle not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            counts[m.group(1)] += 1
    print("== %s (%s)" % (needle, obj))
    print(" ".join("%s:%d" % kv for kv in counts.most_common()))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
