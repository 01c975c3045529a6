"""CPU emulation (numpy / scipy) of the multilevel-preconditioned CG of bfm_b200/csrc/mg.cuh, on the hierarchy the
library itself builds (bfmx_hier_info / bfmx_hier_level): the same aggregates, geometry, scalings, damping rule
and cycle shape.  Test infrastructure: it checks hier.c on the CPU and documents the numerics of the device path."""

from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from bfm_b200 import ext


def hierarchy(lib, mesh):
	"""[dict(n, agg, geom, rowptr, col)] per level; [] when the mesh gets no hierarchy"""

	info = ext.HierInfo()
	assert not lib.lib.bfmx_hier_info(mesh.c_mesh, C.byref(info))

	levels = []

	for l in range(info.n_levels):
		n = info.n_nodes[l]
		last = l == info.n_levels - 1
		agg = np.zeros(n, np.int32)
		geom = np.zeros((n, 2), np.float32)
		rowptr = np.zeros(n + 1, np.int32)

		p = lambda a, t: a.ctypes.data_as(C.POINTER(t))

		assert not lib.lib.bfmx_hier_level(mesh.c_mesh, l, None if last else p(agg, C.c_int32), None if last else p(geom, C.c_float), p(rowptr, C.c_int32), None)

		col = np.zeros(int(rowptr[-1]), np.int32)
		assert not lib.lib.bfmx_hier_level(mesh.c_mesh, l, None, None, p(rowptr, C.c_int32), p(col, C.c_int32))

		levels.append(dict(n=n, agg=agg, geom=geom.astype(np.float64), rowptr=rowptr, col=col))

	return levels


def tentative(level, dofs, sq):
	"""P~ = D^1/2 R of one level as a scipy matrix (dofs per fine node: 2 or 3); sq = sqrt(diag) per fine unknown"""

	n, agg, geom = level["n"], level["agg"].astype(np.int64), level["geom"]
	keep = agg >= 0
	a = np.arange(n)[keep]
	g = agg[keep]
	dx, dy = geom[keep, 0], geom[keep, 1]
	n_coarse = int(agg.max()) + 1

	s = [sq[dofs * a + k] for k in range(dofs)]
	rows = [dofs * a, dofs * a, dofs * a + 1, dofs * a + 1]
	cols = [3 * g, 3 * g + 2, 3 * g + 1, 3 * g + 2]
	vals = [s[0], -dy * s[0], s[1], dx * s[1]]

	if dofs == 3:
		rows.append(3 * a + 2)
		cols.append(3 * g + 2)
		vals.append(s[2])

	return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dofs * n, 3 * n_coarse))


class Emulation:
	def __init__(self, A, levels, gamma=None, omega=1.9, single_precision_p=True, smooth=False, smooth_omega=1.8):
		"""A: the assembled matrix (scipy, DOFs interleaved per node); levels: hierarchy().  smooth: smoothed aggregation
		as k_mg_smooth does it on one GPU - P = (I - w A^) P~ on every level, w = smooth_omega / (largest absolute row
		sum) - with a V-cycle unless gamma says otherwise (the tentative prolongator needs the W-cycle)"""

		gamma = gamma if gamma is not None else (1 if smooth else 2)

		d = np.abs(A.diagonal())
		self.dscale = np.where(d > 0, 1.0 / np.sqrt(np.where(d > 0, d, 1.0)), 1.0)
		self.ops = [(sp.diags(self.dscale) @ A @ sp.diags(self.dscale)).tocsr()]
		self.P = []
		self.gamma = gamma
		self.n_levels = len(levels)

		dsc = self.dscale

		for l in range(len(levels) - 1):
			P = tentative(levels[l], 2 if l == 0 else 3, 1.0 / dsc)

			if smooth:
				bound = float(abs(self.ops[l]).sum(axis=1).max())
				P = (P - (smooth_omega / bound) * (self.ops[l] @ P)).tocsr()

			if l == 0 and single_precision_p:
				P.data = P.data.astype(np.float32).astype(np.float64)

			B = (P.T @ self.ops[l] @ P).tocsr()

			if l + 1 < len(levels) - 1:  # sparse level: scaled to a unit diagonal; the dense last level is not
				dsc = 1.0 / np.sqrt(B.diagonal())
				B = (sp.diags(dsc) @ B @ sp.diags(dsc)).tocsr()
				P = (P @ sp.diags(dsc)).tocsr()

				if l == 0 and single_precision_p:
					P.data = P.data.astype(np.float32).astype(np.float64)

			self.P.append(P.tocsr())
			self.ops.append(B)

		self.omega = [omega / max(1.0, float(abs(M).sum(axis=1).max())) for M in self.ops[:-1]]
		self.dense_inverse = np.linalg.inv(self.ops[-1].toarray())

	def cycle(self, l, g):
		if l == self.n_levels - 1:
			return self.dense_inverse @ g

		A, P, w = self.ops[l], self.P[l], self.omega[l]
		t = g - w * (A @ g)
		z = None

		for visit in range(self.gamma if l >= 1 else 1):
			if visit > 0:
				t = g - A @ z

			mu = self.cycle(l + 1, P.T @ t)
			z = w * g + P @ mu if visit == 0 else z + P @ mu

		return z + w * (g - A @ z)

	def solve(self, b, tol=1e-12, max_iter=2000):
		A = self.ops[0]
		bh = self.dscale * b
		x = np.zeros_like(bh)
		r = bh.copy()
		z = self.cycle(0, r)
		p = z.copy()
		rho = r @ z
		bnorm = np.linalg.norm(bh)

		for it in range(1, max_iter + 1):
			q = A @ p
			alpha = rho / (p @ q)
			x += alpha * p
			r -= alpha * q

			if np.linalg.norm(r) <= tol * bnorm:
				return self.dscale * x, it

			z = self.cycle(0, r)
			rz = r @ z
			p = z + (rz / rho) * p
			rho = rz

		return self.dscale * x, max_iter


def residual_extended(A, b, x):
	"""b - A x with the products and sums in extended precision (x86 long double: what k_residual_dd's double-double
	arithmetic buys on the device), rounded to double at the end"""

	prod = A.data.astype(np.longdouble) * x.astype(np.longdouble)[A.indices]
	sums = np.add.reduceat(prod, A.indptr[:-1])
	sums[np.diff(A.indptr) == 0] = 0

	return (b.astype(np.longdouble) - sums).astype(np.float64)


def solve_refined(emu: Emulation, A, b, mode: str, phi: float, tol=1e-12, max_iter=200, max_events=4):
	"""PCG on the scaled system as solver.cu runs it, with its two ways of getting below FP64's residual floor:

	mode "none"      plain PCG;
	mode "restart"   once r.r <= (phi eps)^2 x^.x^: fold the accumulator into x, recompute the residual of the ORIGINAL
	                 system in extended precision, restart CG on it (solver.cu at phi = 6);
	mode "replace"   the same, but the search direction and rho survive (solver.cu's early replacement at phi = 1e4).

	Returns (x, iterations, events)."""

	eps = np.finfo(np.float64).eps
	Ah, ds = emu.ops[0], emu.dscale
	bh = ds * b
	bnorm = np.linalg.norm(bh)
	base = np.zeros_like(b)
	xh = np.zeros_like(bh)
	r = bh.copy()
	z = emu.cycle(0, r)
	p = z.copy()
	rho = r @ z
	events = 0
	it = 0

	for it in range(1, max_iter + 1):
		q = Ah @ p
		alpha = rho / (p @ q)
		xh += alpha * p
		r -= alpha * q
		rr = r @ r

		if np.sqrt(rr) <= tol * bnorm:
			break

		if mode != "none" and events < max_events and rr <= (phi * eps) ** 2 * (xh @ xh):
			events += 1
			base = base + ds * xh
			xh[:] = 0
			r = ds * residual_extended(A, b, base)

			if mode == "restart":
				z = emu.cycle(0, r)
				p = z.copy()
				rho = r @ z
				continue

		z = emu.cycle(0, r)
		rz = r @ z
		p = z + (rz / rho) * p
		rho = rz

	return base + ds * xh, it, events
