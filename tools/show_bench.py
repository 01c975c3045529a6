"""print the key figures of bench.py JSON lines (stdin or files); ignores non-JSON lines"""
import json, sys, fileinput
for line in fileinput.input():
	line = line.strip()
	if not line.startswith("{"):
		continue
	d = json.loads(line)
	if d.get("impl") == "reference":
		print("reference", round(d["value"]), d["unit"], d["cpu_baseline"]["sample"][:60]); continue
	r = d.get("roofline") or {}
	it = r.get("cg_iteration") or {}
	e = d.get("e2e") or {}
	print(d["config"].get("cells", d["config"]["workload"][:40]), "gpus", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1),
	      "asm", round(d.get("assembly_ms", 0), 2), "solve", round(d.get("solve_ms", 0), 1), "setup", round(d.get("solve_setup_ms", 0), 1),
	      "iters", d.get("cg_iterations"), "coarse", d.get("coarse_dim"), "us/iter", round(it.get("us", 0), 1), "iter_frac", round(it.get("frac", 0) or 0, 3),
	      "spmv_us", round(r.get("us_per_launch", 0), 1), "spmv_frac", round(r.get("frac", 0) or 0, 3), "e2e", round(e.get("value", 0)), "launches", d.get("gpu_launches"), "clocks", (d.get("clocks") or {}).get("sm_mhz"))
