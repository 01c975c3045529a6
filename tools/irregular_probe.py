"""How much does node numbering cost?  The structured plate with its nodes renumbered at random (seeded) - the worst case
for every gather of the solver: q = A p reads p at scattered rows, the aggregates' members are scattered - against the
same plate in natural order.  Prints one JSON line per variant (profiles/r2_summary.md quotes it), and the same randomly numbered plate run on the
internal Morton numbering of bfm_b200/csrc/renumber.c.

    python tools/irregular_probe.py [NXxNY]           # default 1600x400 = 1.28 M nodes, 2.6 M DOF
"""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bfm_b200 import api, ext, workloads  # noqa: E402

cells = sys.argv[1] if len(sys.argv) > 1 else "1600x400"
nx, ny = (int(v) for v in cells.split("x"))
binding = api.default_binding()
assert ext.device_available(binding), binding.lib.bfmx_device_error()


def run(label, coords, elems, renumber=None):
	if renumber is None:
		os.environ.pop("BFM_RENUMBER", None)
	else:
		os.environ["BFM_RENUMBER"] = renumber

	mesh = api.Mesh.from_arrays(coords, elems, binding=binding)
	left = coords[:, 0] == 0.0
	E, nu, rho = workloads.STEEL
	material = api.Material("steel", rho, E, nu, binding=binding)
	rule = api.Rule_gauss_legendre(2, mesh.kind, binding=binding)
	obj = api.Obj(mesh, material, rule)
	instance = api.Instance(obj)

	for kind in (api.Condition.DIRICHLET_X, api.Condition.DIRICHLET_Y):
		cond = api.Condition(mesh, kind, 0.0)
		cond.set_nodes(left)
		instance.add_condition(cond)

	sim = api.Sim(api.CSim.PLANAR_STRESS, binding=binding)
	sim.add_instance(instance)
	force = api.Force_linear(workloads.GRAVITY, binding=binding)
	sim.add_force(force)

	job = ext.Job(sim)
	job.upload()
	job.assemble()
	job.solve()
	job.assemble()
	job.solve()
	s = job.stats()
	spmv_ms = job.spmv_ms(50)
	stored = s["n_slots"] * 36 + (s["n_dofs"] // 2) * 32

	job.download()

	print(json.dumps({
		"numbering": label, "internally_renumbered": ext.internal_numbering(mesh) is not None, "cells": cells, "n_dofs": s["n_dofs"], "cg_iterations": s["cg_iterations"], "mg_levels": s["mg_levels"],
		"solve_ms": s["ms_solve"], "us_per_iteration": (s["ms_solve"] - s["ms_solve_setup"]) * 1e3 / max(s["cg_iterations"], 1),
		"spmv_us": spmv_ms * 1e3, "spmv_stored_format_gbs": stored / (spmv_ms * 1e-3) / 1e9, "assemble_ms": s["ms_assemble"],
	}), flush=True)

	return workloads.effects_view(instance).reshape(-1, 2).copy()


coords, elems = workloads.plate_arrays(nx, ny)
natural = run("natural (row-major)", coords, elems)

rng = np.random.RandomState(0)
perm = rng.permutation(len(coords))          # new id of old node a = perm[a]
inverse = np.empty_like(perm)
inverse[perm] = np.arange(len(perm))

random_coords, random_elems = coords[inverse], perm[elems.astype(np.int64)].astype(np.uint64)

kept = run("random, kept (BFM_RENUMBER=0)", random_coords, random_elems, renumber="0")
auto = run("random, library's choice", random_coords, random_elems)
forced = run("random, Morton (BFM_RENUMBER=1)", random_coords, random_elems, renumber="1")

for label, got in (("kept", kept), ("auto", auto), ("forced", forced)):
	err = np.linalg.norm(got[perm] - natural) / np.linalg.norm(natural)
	print(json.dumps({"variant": label, "same_displacements_rel_l2": float(err)}), flush=True)
