#!/usr/bin/env python
"""Dynamic opcode histogram of a kernel from the ncu source page: warp-level instructions executed per opcode,
optionally divided by a number of warp-frames.  usage: tools/ncu_dyn_hist.py report.ncu-rep [warp_frames]"""
import collections, csv, re, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
hist = collections.Counter(); total = 0
for r in rows[2:]:
    if len(r) != len(hdr): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    if not m: continue
    op = m.group(1)
    key = op.split(".")[0]
    if key in ("IMAD", "LDS", "LEA", "ST", "LD", "STG", "LDG", "SYNCS", "SHF", "ISETP"): key = ".".join(op.split(".")[:2])
    n = int(r[ix["Instructions Executed"]]); hist[key] += n; total += n
print("total warp instructions", total, "" if per is None else f"= {total / per:.2f} per warp-frame")
for k, n in hist.most_common(40):
    print(f"{n:12d} {100.0 * n / total:6.2f}% {'' if per is None else f'{n / per:7.2f}'} {k}")
