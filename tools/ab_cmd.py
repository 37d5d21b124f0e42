#!/usr/bin/env python
"""Same-box A/B of library variants on an arbitrary command that prints 'ms per launch: [...]' (tools/run_config.py).
usage: tools/ab_cmd.py reps "cmd args" variant1 variant2 ..."""
import os, re, subprocess, sys
reps, cmd, variants = int(sys.argv[1]), sys.argv[2].split(), sys.argv[3:]
res = {v: [] for v in variants}
for r in range(reps):
    for v in variants:
        env = dict(os.environ)
        if v != "main":
            env["CRB200_LIB"] = os.path.abspath(f"variants/{v}/libclownresampler_b200.so")
        out = subprocess.run([sys.executable] + cmd, env=env, capture_output=True, text=True)
        m = re.search(r"ms per launch: \[([^\]]*)\]", out.stdout)
        res[v].append(min(float(x) for x in m.group(1).split(",")) if m else out.stderr[-200:])
for v in variants:
    print(v, res[v])
