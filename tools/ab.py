#!/usr/bin/env python
"""Same-box A/B timing of library variants: runs bench.py --no-e2e --no-cpu for each variant, interleaved,
`reps` times and prints ms_per_step per run.  usage: tools/ab.py reps steps variant1 variant2 ..."""
import json
import os
import subprocess
import sys

reps, steps = int(sys.argv[1]), int(sys.argv[2])
variants = sys.argv[3:]
res = {v: [] for v in variants}
for r in range(reps):
    for v in variants:
        env = dict(os.environ)
        if v != "main":
            env["CRB200_LIB"] = os.path.abspath(f"variants/{v}/libclownresampler_b200.so")
        out = subprocess.run([sys.executable, "bench.py", "--steps", str(steps), "--warmup", "3", "--no-e2e", "--no-cpu"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[v].append(round(d["ms_per_step"], 4))
        except Exception:
            res[v].append(out.stderr[-300:])
for v in variants:
    print(v, res[v])
