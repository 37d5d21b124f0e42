#!/usr/bin/env python
"""Opcode histogram of one kernel in a built object / library (cuobjdump -sass): evidence of what the hot loop issues.
usage: tools/sass_hist.py <file.o|.so> <substring of the mangled kernel name> [--full-opcode]"""
import collections, re, subprocess, sys
path, pat = sys.argv[1], sys.argv[2]
full = "--full-opcode" in sys.argv
out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            hist[op if full else op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD", "LDS", "LEA", "STG", "SYNCS", "UBLKCP")) and "." in op else "")] += 1
print(f"# {path} :: *{pat}*  ({sum(hist.values())} instructions, static count over the whole kernel)")
for op, n in hist.most_common():
    print(f"{n:6d} {op}")
