"""Regenerates tests/golden/ from the reference tree (run in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

Inputs copied verbatim (test fixtures, not sources): the meshes and the problem file of BASELINE.json's
configs 1-2, and the reference's own golden vectors data/U.txt, data/V.txt.  Derived inputs: a
full-format copy of gear60 (the file shipped by the reference lacks the edge and domain sections its own
reader requires) with three named boundary domains, and mixed-BC problem files.

Outputs: ref_outputs.npz - displacements, right-hand side, RCM permutation, bandwidths and numeric
pattern digests produced by the UNMODIFIED reference library (oracle/_ref/libbfm_ref.so) for every case
in CASES.  The GPU box has no reference tree: tests read these files only.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import sys
from collections import Counter

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BFM_REFERENCE", "/root/reference")

sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def copy_inputs():
	os.makedirs(os.path.join(HERE, "meshes"), exist_ok=True)
	os.makedirs(os.path.join(HERE, "problems"), exist_ok=True)

	for name in ("8.lepl1110", "bridge.obj", "bridge-dam.obj"):
		shutil.copyfile(os.path.join(REF, "meshes", name), os.path.join(HERE, "meshes", name))

	shutil.copyfile(os.path.join(REF, "problems", "problem.txt"), os.path.join(HERE, "problems", "problem.txt"))
	shutil.copyfile(os.path.join(REF, "data", "U.txt"), os.path.join(HERE, "lepl8_U.txt"))
	shutil.copyfile(os.path.join(REF, "data", "V.txt"), os.path.join(HERE, "lepl8_V.txt"))

	for name in ("8.lepl1110", "bridge.obj", "bridge-dam.obj"):
		os.chmod(os.path.join(HERE, "meshes", name), 0o644)

	for path in ("problems/problem.txt", "lepl8_U.txt", "lepl8_V.txt"):
		os.chmod(os.path.join(HERE, path), 0o644)


def convert_gear60():
	"""gear60.lepl1110 -> full LEPL1110 format with boundary edges + domains Clamp / Load / Free"""

	lines = open(os.path.join(REF, "meshes", "gear60.lepl1110")).read().split("\n")
	n = int(lines[0].split()[3])
	node_lines = lines[1:1 + n]
	xy = np.array([[float(v) for v in line.split(":")[1].split()] for line in node_lines])
	nt = int(lines[1 + n].split()[3])
	tri_lines = lines[2 + n:2 + n + nt]
	tri = np.array([[int(v) for v in line.split(":")[1].split()] for line in tri_lines])

	count = Counter()
	owner = {}

	for e, (a, b, c) in enumerate(tri):
		for u, v in ((a, b), (b, c), (c, a)):
			key = (min(u, v), max(u, v))
			count[key] += 1
			owner[key] = (e, u, v)

	boundary = sorted((owner[k] for k, v in count.items() if v == 1))
	mid_y = np.array([(xy[u, 1] + xy[v, 1]) / 2 for (_, u, v) in boundary])

	domains = {
		"Clamp": [i for i in range(len(boundary)) if mid_y[i] < -200],
		"Load": [i for i in range(len(boundary)) if mid_y[i] > 200],
		"Free": [i for i in range(len(boundary)) if -200 <= mid_y[i] <= 200],
	}

	out = [f"Number of nodes {n} "] + node_lines
	out.append(f"Number of edges {len(boundary)} ")
	out += [f"{e:6d} : {u:6d} {v:6d} " for (e, u, v) in boundary]
	out.append(f"Number of triangles {nt} ")
	out += tri_lines
	out.append(f"Number of domains {len(domains)}")

	for i, (name, ids) in enumerate(domains.items()):
		out.append(f"  Domain : {i:6d} ")
		out.append(f"  Name : {name}")
		out.append(f"  Number of elements : {len(ids):6d}")

		for s in range(0, len(ids), 10):
			out.append("".join(f"{v:6d}" for v in ids[s:s + 10]))

	with open(os.path.join(HERE, "meshes", "gear60_full.lepl1110"), "w") as f:
		f.write("\n".join(out) + "\n")


PROBLEMS = {
	# config 3: mixed Dirichlet / Neumann on the gear
	"gear60_mixed.txt": """Type of problem    :  Planar stresses
Young modulus      :  2.1100000e+11
Poisson ratio      :  3.0000000e-01
Mass density       :  7.8500000e+03
Gravity            :  9.8100000e+00
Boundary condition :  Dirichlet-X        =  0.0000000e+00 : Clamp
Boundary condition :  Dirichlet-Y        =  0.0000000e+00 : Clamp
Boundary condition :  Neumann-Y          = -1.0000000e+06 : Load
Boundary condition :  Neumann-Tangent    =  2.0000000e+05 : Load
""",
	# every condition kind the planar path knows, with non-zero Dirichlet values, on config 1's mesh.
	# NB the mesh file names these domains "Entity N " WITH a trailing blank and the parser compares the
	# rest of the line verbatim (reference ez.c:113-160), so the lines below end in a blank too.
	"lepl8_all_kinds.txt": """Type of problem    :  Planar stresses
Young modulus      :  2.1100000e+11
Poisson ratio      :  3.0000000e-01
Mass density       :  7.8500000e+03
Gravity            :  9.8100000e+00
Boundary condition :  Dirichlet-X        =  1.0000000e-04 : Symmetry
Boundary condition :  Neumann-X          =  3.0000000e+05 : Entity 2 
Boundary condition :  Dirichlet-Y        = -2.0000000e-04 : Bottom
Boundary condition :  Neumann-Normal     =  1.0000000e+05 : Entity 4 
Boundary condition :  Neumann-Tangent    = -5.0000000e+04 : Entity 5 
Boundary condition :  Dirichlet-Tangent  =  3.0000000e-05 : Entity 6 
Boundary condition :  Neumann-Y          = -7.0000000e+05 : Bottom
Boundary condition :  Dirichlet-Normal   = -1.0000000e-05 : Entity 1 
""",
	# the axisymmetric path with its Dirichlet / Neumann-X/Y branches
	"lepl8_axisym.txt": """Type of problem    :  Axi-symetric problem
Young modulus      :  2.1100000e+11
Poisson ratio      :  3.0000000e-01
Mass density       :  7.8500000e+03
Gravity            :  9.8100000e+00
Boundary condition :  Dirichlet-X        =  0.0000000e+00 : Symmetry
Boundary condition :  Dirichlet-Y        =  0.0000000e+00 : Bottom
Boundary condition :  Neumann-X          =  2.0000000e+05 : Entity 3 
""",
}


def pattern_digest(A: np.ndarray) -> str:
	"""sha1 of the numeric non-zero pattern (row-major boolean mask, packed)"""

	return hashlib.sha1(np.packbits(A != 0).tobytes()).hexdigest()


def main():
	copy_inputs()
	convert_gear60()

	for name, text in PROBLEMS.items():
		with open(os.path.join(HERE, "problems", name), "w") as f:
			f.write(text)

	from oracle import ref

	import cases

	binding = ref.binding()
	out = {}

	only = set(sys.argv[1:])
	target = os.path.join(HERE, "ref_outputs.npz")

	if only and os.path.exists(target):
		out = {k: v for k, v in np.load(target).items() if k.split("/")[0] not in only}

	for name in cases.CASES:
		if only and name not in only:
			continue

		case = cases.build(name, binding)
		sim = case.sim

		sim.run()
		out[f"{name}/effects"] = case.instance.effects.copy()

		from bfm_b200.api import System

		system = System(sim, case.instance)
		out[f"{name}/b"] = system.b()

		A = system.dense(copy=cases.CASES[name])  # FULL row-major: the reference's n*n array (a view for the heavy cases)
		out[f"{name}/nnz"] = np.array(int(np.count_nonzero(A)))
		out[f"{name}/bandwidth_natural"] = np.array(system.bandwidth())

		if cases.CASES[name]:
			out[f"{name}/pattern_sha1"] = np.array(pattern_digest(A))
			out[f"{name}/abs_sum"] = np.array(np.abs(A).sum())

		del A

		system.renumber()
		perm, inv_perm = system.perm()
		out[f"{name}/perm"] = perm.astype(np.int64)
		out[f"{name}/bandwidth_rcm"] = np.array(system.bandwidth())

		print(f"{name}: n = {system.n}, nnz = {out[f'{name}/nnz']}, bandwidth {out[f'{name}/bandwidth_natural']} -> {out[f'{name}/bandwidth_rcm']}", flush=True)

	np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)


if __name__ == "__main__":
	main()
