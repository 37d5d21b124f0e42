#!/usr/bin/env python
"""Same-box A/B of the host bulk path (e2e): usage tools/ab_e2e.py variant..."""
import json, os, subprocess, sys
for v in sys.argv[1:]:
    env = dict(os.environ)
    if v != "main":
        env["CRB200_LIB"] = os.path.abspath(f"variants/{v}/libclownresampler_b200.so")
    out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "1", "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(v, round(d["e2e"]["s_per_step"], 4), round(d["e2e"]["value"]))
    except Exception:
        print(v, out.stderr[-400:])
