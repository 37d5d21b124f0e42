#!/usr/bin/env python
"""Turns the raw evidence of a measurement pass (gpurun_out/r02/, written on the GPU box by tools/gather_r02.sh and the 2- / 8-GPU bench
commands in profiles/r02_README.txt) into the files committed under profiles/: bench lines, ncu summaries with dynamic opcode histograms,
DRAM traffic, launch shares, SASS histograms of the built library, microbenchmarks.  Runs on the CPU box (ncu -i reads the reports)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, G = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out", "r02")
R = "r02_"


def copy(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, R + dst))


def rows_of(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    return rows[0], rows[1:]


# bench lines (bench_n1, configs, reference arm, the 2- and 8-GPU runs of tools/bench_n2.sh / bench_n8.sh, PCIe probes) are copied
# into profiles/ by hand from the run they belong to: several boxes and builds contribute over a round (profiles/r02_README.txt)
for src, dst in [("warp_time.txt", "warp_time.txt"), ("pipe_microbench.jsonl", "pipe_microbench.jsonl"), ("overlap.jsonl", "overlap.jsonl"),
                 ("launches.csv", "launches.csv"), ("traffic.csv", "traffic.csv")]:
    copy(src, dst)

# DRAM traffic of the full-workload launch
hdr, rows = rows_of(os.path.join(G, "traffic.csv"))
vals = {r[hdr.index("Metric Name")]: int(r[hdr.index("Metric Value")].replace(",", "")) for r in rows}
bench = json.load(open(os.path.join(G, "bench_n1.json")))
t = {"command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:crb_tiled -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu",
     "kernel": rows[0][hdr.index("Kernel Name")], "workload": bench["config"]["workload"] + " (full bench workload, one launch)",
     "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
     "dram_bytes_per_launch_full_workload": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
     "algorithmic_bytes_per_launch": bench["roofline"]["algorithmic_bytes_per_launch"], "gpu_time_duration_ns_under_ncu": vals["gpu__time_duration.sum"]}
json.dump(t, open(os.path.join(P, R + "traffic.json"), "w"), indent=1)

# time share per kernel of the bench command
hdr, rows = rows_of(os.path.join(G, "launches.csv"))
kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows:
    if r[mn] == "gpu__time_duration.sum":
        agg[r[kn]][0] += float(r[mv].replace(",", ""))
        agg[r[kn]][1] += 1
tot = sum(v[0] for v in agg.values())
with open(os.path.join(P, R + "launch_shares.txt"), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu` (profiles/r02_launches.csv): time share per kernel\n"
            "# (cold-cache, serialised launches; compare shares, not absolutes).  The timed region of bench.py contains only crb_tiled_kernel launches;\n"
            "# crb_noise_kernel and the torch fill are input set-up outside the timed region.\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{v[0] / 1e6:10.3f} ms {v[1]:5d} launches {100 * v[0] / tot:5.1f}%  {k[:110]}\n")

# ncu summaries + dynamic opcode histograms + top stall sites
warp_frames = {"stereo": 16 * 11520039 // 32, "mono": 1024 * 480010 // 32, "8ch": 300 * 44100 // 32, "sk": 16 * 2646000 // 32}
what = {"stereo": "python bench.py --steps 2 --warmup 3 --streams 16 --seconds 240 --no-e2e --no-cpu   (config 2's kernel, 16 streams x 240 s)",
        "mono": "python bench.py --config 4 --steps 2 --warmup 3 --no-e2e --no-cpu   (config 4 as one bulk launch: 1024 mono voices x 10 s)",
        "8ch": "python tools/run_config.py 8 192000 44100 300 1 3   (config 3's kernel, 300 s of the hour)",
        "sk": "python tools/run_config.py 2 48000 44100 60 16 3   (slightly stretched kernel, stereo 48 -> 44.1 kHz)"}
for name in ("stereo", "mono", "8ch", "sk"):
    rep = os.path.join(G, f"ncu_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = f"# ncu --set full --clock-control none --import-source on -k regex:crb_tiled ... {what[name]}\n"
    out += subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    out += "\n# dynamic opcode histogram (warp-level instructions executed, ncu source page); last column: per warp-frame (32 output frames)\n"
    out += "\n".join(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_dyn_hist.py"), rep, str(warp_frames[name])], capture_output=True, text=True).stdout.splitlines()[:32])
    out += "\n\n# top stall sites (warp-state samples per SASS line)\n"
    out += "\n".join(l[:200] for l in subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_top.py"), rep, "10"], capture_output=True, text=True).stdout.splitlines())
    open(os.path.join(P, R + f"ncu_{name}_summary.txt"), "w").write(out + "\n")

# static SASS histograms of the built objects (what the hot kernels are made of: UBLKCP = TMA bulk copy, SYNCS = mbarrier)
for name, obj, pat in (("stereo", "crb_inst_k1_p0.o", "ILi2ELi1ELi1E"), ("mono", "crb_inst_k1_p0.o", "ILi1ELi1ELi1E"), ("8ch", "crb_inst_k0_p1.o", "ILi8ELi1ELi0E"), ("sk_stereo6", "crb_inst_k6_p0.o", "ILi2ELi1ELi6E")):
    o = os.path.join(ROOT, "build", "obj", obj)
    if os.path.exists(o):
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_hist.py"), o, pat, "--full-opcode"], capture_output=True, text=True).stdout
        open(os.path.join(P, R + f"sass_hist_{name}.txt"), "w").write(txt)
print(json.dumps(t))
print(open(os.path.join(P, R + "launch_shares.txt")).read())
