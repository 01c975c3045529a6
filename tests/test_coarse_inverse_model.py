"""numpy model of the coarse operator's inversion (bfm_b200/csrc/coarse.cuh: coarse_invert and the k_gj_* kernels).

It pins the ALGORITHM the kernels implement - the CUDA code itself is covered by the GPU parity tests through the
iteration counts and results it produces:
  - in-place blocked Gauss-Jordan on an SPD matrix without pivoting (block k: P = E_KK^-1; E_K* = P E_K*;
    E_ij -= E_iK E_Kj; E_*K = -E_*K P; E_KK = P);
  - the band skip: E starts banded (adjacent aggregates have close numbers) and after pivot blocks 0..k everything
    at or beyond lim = 32 (k + 1) + half_bw is still zero in both panels, so those tiles are never touched;
  - the row distribution: rank r only updates its own row blocks, with the pivot row panel and P coming from the
    block's owner - and ends up holding exactly its rows of E^-1.
"""

import numpy as np
import pytest

BLOCK = 32


def banded_spd(n, half_bw, seed=0):
	rng = np.random.default_rng(seed)
	A = np.zeros((n, n))

	for i in range(n):
		lo = max(0, i - half_bw)
		A[i, lo:i] = rng.uniform(-1, 1, i - lo)

	A = A + A.T
	A += np.diag(np.abs(A).sum(axis=1) + 1.0)
	return A


def invert(E, half_bw, world=1):
	"""returns the per-rank matrices after the distributed in-place inversion (list of n x n arrays; only the rows a
	rank owns are meaningful) and the number of 32 x 32 tile updates performed"""

	n = E.shape[0]
	blocks = n // BLOCK
	per = -(-blocks // world)
	own = [(min(r * per, blocks) * BLOCK, min((r + 1) * per, blocks) * BLOCK) for r in range(world)]
	local = [E.copy() for _ in range(world)]  # every rank probes the whole of E; it then only maintains its rows
	tiles = 0

	for k in range(blocks):
		K = slice(k * BLOCK, (k + 1) * BLOCK)
		lim = min(n, (k + 1) * BLOCK + half_bw)
		owner = k // per

		# owner: k_gj_diag, k_gj_row (columns below lim only), k_gj_bcast
		Eo = local[owner]
		P = np.linalg.inv(Eo[K, K])
		R = Eo[K, :].copy()
		R[:, :lim] = P @ Eo[K, :lim]
		R[:, K] = Eo[K, K]  # the pivot block itself is not multiplied (J == k is skipped)
		Eo[K, :] = R

		for r in range(world):
			lo, hi = own[r]
			Er = local[r]

			for i0 in range(lo, min(hi, lim), BLOCK):  # k_gj_update: tiles at or beyond lim are skipped
				I = slice(i0, i0 + BLOCK)

				if i0 == k * BLOCK:
					continue

				C = Er[I, K].copy()

				for j0 in range(0, lim, BLOCK):
					if j0 == k * BLOCK:
						continue

					J = slice(j0, j0 + BLOCK)
					Er[I, J] -= C @ R[:, J]
					tiles += 1

				Er[I, K] = -C @ P  # k_gj_col

			if lo <= k * BLOCK < hi:
				Er[K, K] = P

	return local, own, tiles


@pytest.mark.parametrize("n,half_bw,world", [(256, 40, 1), (256, 40, 3), (320, 95, 2), (192, 191, 4), (256, 10, 8)])
def test_block_gauss_jordan_with_band_skip_and_row_distribution(n, half_bw, world):
	E = banded_spd(n, half_bw)
	want = np.linalg.inv(E)

	local, own, tiles = invert(E, half_bw, world)

	for r, (lo, hi) in enumerate(own):
		assert np.allclose(local[r][lo:hi], want[lo:hi], rtol=1e-9, atol=1e-12 * np.abs(want).max()), r

	# the skip is worth it: far fewer tile updates than the dense blocks * (blocks - 1)^2
	blocks = n // BLOCK

	if half_bw < n // 4:
		assert tiles < 0.7 * blocks * (blocks - 1) ** 2
