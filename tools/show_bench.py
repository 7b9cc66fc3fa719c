import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "it/s | e2e", d["e2e"] and round(d["e2e"]["value"], 2), "| ms/step", round(d["ms_per_step"], 3), "| launches", d["gpu_launches"])
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items() if k in ("achieved", "frac", "launch_ms")})
if "roofline_dense" in d:
    print("dense", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["roofline_dense"].items() if k in ("achieved", "frac", "launch_ms")})
    print("isolated", {k: round(v, 4) for k, v in d["phase_ms_isolated"].items()})
print("per_solve", {k: round(v, 3) for k, v in d["phase_ms_per_solve"].items()}, "clocks", d["clocks"])
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k in ("roofline_scaled",):
    if k in d: print(k, d[k])
