#!/usr/bin/env python
"""Same-box A/B of library variants on `bench.py --config C --no-e2e --no-cpu`: prints ms_per_step per run.
usage: tools/ab_config.py config reps steps variant1 variant2 ...
("main" = the in-tree build, "NAME=VALUE" = the in-tree build with that environment variable, else variants/<name>/)"""
import json, os, subprocess, sys
cfg, reps, steps, variants = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4:]
res = {v: [] for v in variants}
for r in range(reps):
    for v in variants:
        env = dict(os.environ)
        if "=" in v:
            env[v.split("=", 1)[0]] = v.split("=", 1)[1]
        elif v != "main":
            env["CRB200_LIB"] = os.path.abspath(f"variants/{v}/libclownresampler_b200.so")
        out = subprocess.run([sys.executable, "bench.py", "--config", cfg, "--steps", str(steps), "--warmup", "3", "--no-e2e", "--no-cpu"], env=env, capture_output=True, text=True)
        try:
            line = json.loads(out.stdout.strip().splitlines()[-1])
            res[v].append(round(line["ms_per_step"], 4))
            if r == 0 and isinstance(line["config"].get("cases"), list):
                print(cfg, v, "cases", json.dumps(line["config"]["cases"]))
        except Exception:
            res[v].append(out.stderr[-300:])
for v in variants:
    print(cfg, v, res[v])
