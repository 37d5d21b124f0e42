import json, os, subprocess, sys
for v in sys.argv[1:]:
    env = dict(os.environ)
    if v != "main":
        env["CRB200_LIB"] = os.path.abspath(f"variants/{v}/libclownresampler_b200.so")
    out = subprocess.run([sys.executable, "bench.py"], env=env, capture_output=True, text=True)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print(v, "e2e s/step", round(d["e2e"]["s_per_step"], 4), round(d["e2e"]["value"]), "kernel ms", round(d["ms_per_step"], 3), "cpu", round(d["cpu_baseline"]["value"], 1))
