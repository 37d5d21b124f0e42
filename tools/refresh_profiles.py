#!/usr/bin/env python
"""Regenerates the derived files under profiles/ from the raw ncu csv logs of the last measurement pass:
r01_traffic.json (DRAM bytes of the bench kernel) and r01_launch_shares.txt (time share per kernel)."""
import collections
import csv
import json
import os

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def rows_of(name):
    rows = [r for r in csv.reader(l for l in open(os.path.join(P, name)) if l.startswith('"'))]
    return rows[0], rows[1:]


hdr, rows = rows_of("r01_traffic.csv")
vals = {r[hdr.index("Metric Name")]: int(r[hdr.index("Metric Value")].replace(",", "")) for r in rows}
t = json.load(open(os.path.join(P, "r01_traffic.json")))
t["dram_bytes_read"], t["dram_bytes_write"] = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
t["dram_bytes_per_launch_full_workload"] = t["dram_bytes_read"] + t["dram_bytes_write"]
t["gpu_time_duration_ns_under_ncu"] = vals["gpu__time_duration.sum"]
json.dump(t, open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)

hdr, rows = rows_of("r01_launches.csv")
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    if r[mn] == "gpu__time_duration.sum":
        agg[r[kn]][0] += float(r[mv].replace(",", ""))
        agg[r[kn]][1] += 1
tot = sum(v[0] for v in agg.values())
with open(os.path.join(P, "r01_launch_shares.txt"), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu` (profiles/r01_launches.csv): time share per kernel\n"
            "# (cold-cache, serialised launches; compare shares, not absolutes).  The timed region of bench.py contains only crb_tiled_kernel launches;\n"
            "# crb_noise_kernel and the torch fill are input set-up outside the timed region.\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{v[0] / 1e6:10.3f} ms {v[1]:5d} launches {100 * v[0] / tot:5.1f}%  {k[:110]}\n")
print(json.dumps(t))
print(open(os.path.join(P, "r01_launch_shares.txt")).read())
