"""numpy emulation of the row-partitioned Jacobi-PCG (bfm_b200/csrc/solver.cu + dist.cu) on top of the
HOST partition the library computes (bfmx_partition_*): each emulated rank only ever reads its owned
matrix rows, its local (owned + ghost) vector entries and what the halo plan delivers.  Used by
tests/test_partition.py serially (all ranks in one process) and over gloo (one process per rank)."""

from __future__ import annotations

import numpy as np


class RankView:
	"""what one rank holds: owned rows of the BC'd system in LOCAL column numbering"""

	def __init__(self, part: dict, A, b):
		self.part = part
		self.l2g = part["l2g"].astype(np.int64)
		self.own = slice(int(part["own_begin"]), int(part["own_end"]))
		self.n_local = len(self.l2g)

		g2l = {int(g): l for l, g in enumerate(self.l2g)}
		own_nodes = self.l2g[self.own]
		rows = np.stack([2 * own_nodes, 2 * own_nodes + 1], axis=1).reshape(-1)
		sub = A[rows].tocsr()

		# every column an owned row touches must be a local node: that is the halo plan's promise
		cols = sub.indices
		local_cols = np.array([2 * g2l[int(c) // 2] + int(c) % 2 for c in cols], dtype=np.int64)  # KeyError = broken partition

		import scipy.sparse as sp

		self.A = sp.csr_matrix((sub.data, local_cols, sub.indptr), shape=(len(rows), 2 * self.n_local))
		self.b = b[rows]
		self.rows = rows
		self.diag = A.diagonal()[rows]

	def dofs(self, nodes):
		nodes = np.asarray(nodes, dtype=np.int64)
		return np.stack([2 * nodes, 2 * nodes + 1], axis=1).reshape(-1)

	def pack(self, v, i):
		"""entries of local vector v that neighbour i ghosts"""

		p = self.part
		idx = p["send_idx"][p["send_ptr"][i]:p["send_ptr"][i + 1]]
		return v[self.dofs(idx)].copy()

	def unpack(self, v, i, data):
		p = self.part
		beg, cnt = int(p["recv_begin"][i]), int(p["recv_count"][i])
		v[2 * beg:2 * (beg + cnt)] = data


def pcg(views, exchange, allsum, tol=1e-12, max_iter=20000):
	"""views: the RankViews this process emulates; exchange(vectors) refreshes ghosts of one local vector per
	view; allsum(list of partials) returns the global sum.  Returns the owned solution blocks."""

	V = views
	own = [slice(2 * v.own.start, 2 * v.own.stop) for v in V]

	scale = []

	for v in V:
		s = np.zeros(2 * v.n_local)
		d = np.abs(v.diag)
		s[own[V.index(v)]] = np.where(d > 0, 1 / np.sqrt(np.where(d > 0, d, 1)), 1.0)
		scale.append(s)

	exchange(scale)

	def spmv(i, p):
		v = V[i]
		return scale[i][own[i]] * (v.A @ (scale[i] * p))

	bh = [scale[i][own[i]] * V[i].b for i in range(len(V))]
	x = [np.zeros_like(b) for b in bh]
	r = [b.copy() for b in bh]
	p = [np.zeros(2 * v.n_local) for v in V]

	for i in range(len(V)):
		p[i][own[i]] = r[i]

	rho = allsum([float(ri @ ri) for ri in r])
	bnorm2 = rho

	for it in range(max_iter):
		exchange(p)
		q = [spmv(i, p[i]) for i in range(len(V))]
		alpha = rho / allsum([float(p[i][own[i]] @ q[i]) for i in range(len(V))])

		for i in range(len(V)):
			x[i] += alpha * p[i][own[i]]
			r[i] -= alpha * q[i]

		new = allsum([float(ri @ ri) for ri in r])

		if new <= tol * tol * bnorm2:
			break

		beta = new / rho
		rho = new

		for i in range(len(V)):
			p[i][own[i]] = r[i] + beta * p[i][own[i]]

	return [scale[i][own[i]] * x[i] for i in range(len(V))], it + 1
