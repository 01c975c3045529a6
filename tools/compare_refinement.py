"""The two ways of getting below FP64's residual floor (bfm_b200/csrc/solver.cu), side by side on one plate: the restart
at the floor (BFM_CG_REPLACE=0) and the early residual replacement that keeps the search direction (the default on one
GPU).  Prints iterations, times and the relative L2 difference of the two displacement fields.

    python tools/compare_refinement.py [NXxNY]
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bfm_b200 import api, ext, workloads  # noqa: E402

cells = sys.argv[1] if len(sys.argv) > 1 else "10000x2500"
nx, ny = (int(v) for v in cells.split("x"))
binding = api.default_binding()
assert ext.device_available(binding), binding.lib.bfmx_device_error()

case = workloads.plate_case(nx, ny, binding=binding)
fields = {}

for label, env in (("restart at the floor", "0"), ("early replacement", "1"), ("no refinement", None)):
	if env is None:
		os.environ["BFM_CG_REFINE"] = "0"
	else:
		os.environ["BFM_CG_REPLACE"] = env

	for _ in range(2):
		case.sim.run()

	s = ext.last_stats(binding)
	fields[label] = workloads.effects_view(case.instance).copy()

	print(json.dumps({
		"variant": label, "cells": cells, "n_dofs": s["n_dofs"], "cg_iterations": s["cg_iterations"], "refinement_events": s["cg_restarts"], "solve_ms": s["ms_solve"],
		"cg_rel_residual": s["cg_rel_residual"], "cg_true_rel_residual": s["cg_true_rel_residual"], "cg_backward_error": s["cg_backward_error"],
	}), flush=True)

base = fields["restart at the floor"]

for label in ("early replacement", "no refinement"):
	print(json.dumps({"variant": label, "rel_l2_vs_restart_at_the_floor": float(np.linalg.norm(fields[label] - base) / np.linalg.norm(base))}), flush=True)
