import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "launches_per_step")}, "e2e", round(d["e2e"]["value"]), "conv frac", round(d["roofline"]["frac"], 4),
      "us", round(d["roofline"]["us_per_launch"], 1), "step tensor frac", round(d["step_tensor_roofline"]["frac"], 4), "clocks", d["clocks"], "cpu", round(d["cpu_baseline"]["value"]))
