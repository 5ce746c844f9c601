#!/usr/bin/env python3
"""Static SASS opcode mix per kernel: python tools/sass_mix.py <lib.so> <substring of kernel name> [top]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, mix = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); mix[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur: mix[cur][m.group(1)] += 1
for fn, c in mix.items():
    if pat in fn:
        print(fn, sum(c.values()))
        for op, n in c.most_common(top): print("   %6d %s" % (n, op))
