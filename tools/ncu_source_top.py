#!/usr/bin/env python
"""Top stall sites of a kernel from the ncu source page (SASS view).
Usage: tools/ncu_source_top.py prof.ncu-rep [N]"""
import csv
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[2:]
    total = sum(int(r[ix["# Samples"]]) for r in body if len(r) == len(hdr))
    print("total samples", total)
    agg = {c: sum(int(r[ix[c]]) for r in body if len(r) == len(hdr)) for c in stall_cols}
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ranked = sorted((r for r in body if len(r) == len(hdr)), key=lambda r: -int(r[ix["# Samples"]]))[:top]
    for r in ranked:
        st = {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]])}
        print(f'{r[ix["# Samples"]]:>6s} {r[ix["Source"]].strip():60s} {st}')


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
