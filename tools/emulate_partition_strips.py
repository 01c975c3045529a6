"""Where do the extra iterations on several GPUs come from?  numpy twin of the smoothed-aggregation V-cycle
(tests/mg_emulation.py) on the library's own hierarchy of a plate cut into horizontal bands: on a partitioned job the
nodes next to a cut keep the tentative prolongator (hier.c: build_transfer), i.e. the smoothing (I - w A^) is masked
out on strips along the cuts.  The script masks it on chosen levels and counts PCG iterations.  CPU only; the numbers
quoted in DESIGN.md section 5 and profiles/r2_summary.md come from `python tools/emulate_partition_strips.py 2800 700`
(3.9 M DOF, ~10 minutes).

    python tools/emulate_partition_strips.py [NX NY]
"""

import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

os.environ["BFM_MG_SMOOTH"] = "1"

import cases  # noqa: E402
import mg_emulation  # noqa: E402
from bfm_b200 import api  # noqa: E402


class Masked(mg_emulation.Emulation):
	"""smoothed aggregation with the smoothing switched off on the rows masks[l] marks False; gammas[l] visits of
	level l + 1 from level l"""

	def __init__(self, A, levels, masks=None, gammas=None, omega=1.9, smooth_omega=1.8):
		d = np.abs(A.diagonal())
		self.dscale = np.where(d > 0, 1.0 / np.sqrt(np.where(d > 0, d, 1.0)), 1.0)
		self.ops = [(sp.diags(self.dscale) @ A @ sp.diags(self.dscale)).tocsr()]
		self.P = []
		self.n_levels = len(levels)
		self.gammas = gammas or [1] * len(levels)
		dsc = self.dscale

		for l in range(len(levels) - 1):
			dofs = 2 if l == 0 else 3
			P = mg_emulation.tentative(levels[l], dofs, 1.0 / dsc)
			bound = float(abs(self.ops[l]).sum(axis=1).max())
			S = (smooth_omega / bound) * (self.ops[l] @ P)

			if masks is not None and masks[l] is not None:
				S = sp.diags(np.repeat(masks[l], dofs).astype(np.float64)) @ S

			P = (P - S).tocsr()
			B = (P.T @ self.ops[l] @ P).tocsr()

			if l + 1 < len(levels) - 1:
				dsc = 1.0 / np.sqrt(B.diagonal())
				B = (sp.diags(dsc) @ B @ sp.diags(dsc)).tocsr()
				P = (P @ sp.diags(dsc)).tocsr()

			self.P.append(P)
			self.ops.append(B)

		self.omega = [omega / max(1.0, float(abs(M).sum(axis=1).max())) for M in self.ops[:-1]]
		self.dense_inverse = np.linalg.inv(self.ops[-1].toarray())

	def cycle(self, l, g):
		if l == self.n_levels - 1:
			return self.dense_inverse @ g

		A, P, w = self.ops[l], self.P[l], self.omega[l]
		t = g - w * (A @ g)
		z = None

		for visit in range(self.gammas[l]):
			if visit > 0:
				t = g - A @ z

			mu = self.cycle(l + 1, P.T @ t)
			z = w * g + P @ mu if visit == 0 else z + P @ mu

		return z + w * (g - A @ z)


def strip_masks(levels, coords, parts_x, parts_y):
	"""per level: True where a node has no neighbour in another part.  Parts are the cells of a parts_x x parts_y grid
	over the plate: (1, N) = horizontal bands, what the row partition of the row-major plate gives; (4, 2) = compact
	blocks, what a partition of the Morton numbering would give.  The reference point of an aggregate is the centroid
	of its members"""

	masks = []
	x = coords[:, 0] / coords[:, 0].max()
	y = coords[:, 1] / coords[:, 1].max()

	for l, L in enumerate(levels[:-1]):
		n = L["n"]
		part = np.minimum((y * parts_y).astype(int), parts_y - 1) * parts_x + np.minimum((x * parts_x).astype(int), parts_x - 1)
		rows = np.repeat(np.arange(n), np.diff(L["rowptr"]))
		cut = np.zeros(n, bool)
		cut[rows[part[rows] != part[L["col"]]]] = True
		masks.append(~cut)

		agg = L["agg"]
		members = agg >= 0
		count = np.maximum(np.bincount(agg[members], minlength=levels[l + 1]["n"]), 1)
		x = np.bincount(agg[members], weights=x[members], minlength=levels[l + 1]["n"]) / count
		y = np.bincount(agg[members], weights=y[members], minlength=levels[l + 1]["n"]) / count

	return masks


def main():
	nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1600, 400)
	lib = api.default_binding()
	case = cases.build(f"plate_{nx}x{ny}", lib)
	levels = mg_emulation.hierarchy(lib, case.mesh)
	system = cases.oracle_problem(case).system()
	A, b = system.scipy().tocsr(), system.b.copy()
	coords = case.mesh.coords_array
	n_sparse = len(levels) - 1

	print(f"plate {nx}x{ny}: {A.shape[0]} DOF, levels {[L['n'] for L in levels]}", flush=True)

	def only(masks, upto):
		return [m if l < upto else None for l, m in enumerate(masks)]

	two, eight, blocks = strip_masks(levels, coords, 1, 2), strip_masks(levels, coords, 1, 8), strip_masks(levels, coords, 4, 2)

	variants = [
		("one part", None, None),
		("2 bands, strip on the mesh level only", only(two, 1), None),
		("2 bands, strips on levels 0-1", only(two, 2), None),
		("2 bands, strips on every level", two, None),
		("8 bands, strips on every level", eight, None),
		("8 blocks (4 x 2), strips on every level", blocks, None),
		("8 bands, every level; level 1 visits level 2 twice", eight, [1, 2] + [1] * n_sparse),
		("8 bands, every level; level 0 visits level 1 twice", eight, [2] + [1] * n_sparse),
	]

	for label, masks, gammas in variants:
		t0 = time.time()
		emu = Masked(A, levels, masks=masks, gammas=gammas)
		_, iterations = emu.solve(b, tol=1e-12)
		print(f"{label:55s} {iterations:4d} iterations  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
	main()
