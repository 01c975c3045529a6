"""Pins the oracle: the C restatement (oracle/bfm_oracle.c) must reproduce, bit for bit,
  (1) the reference's own golden vectors data/U.txt, data/V.txt (tests/golden/lepl8_{U,V}.txt),
  (2) the outputs the unmodified reference library produced for every case (tests/golden/ref_outputs.npz,
      written by tests/golden/make_golden.py),
  (3) the reference library itself, live, when oracle/_ref/libbfm_ref.so is present.
"""

import hashlib
import os

import numpy as np
import pytest

import cases


def _format_uv(values):
	"""reference ez.c:211-229"""

	out = f"Number of nodes {len(values)}\n"

	for i, v in enumerate(values):
		out += "%14.7e" % v

		if (i + 1) % 3 == 0:
			out += "\n"

	return out + "\n"


def test_reference_golden_vectors():
	case = cases.build_oracle_only("lepl8")
	x = case.run()

	assert _format_uv(x[:, 0]) == open(os.path.join(cases.GOLDEN, "lepl8_U.txt")).read()
	assert _format_uv(x[:, 1]) == open(os.path.join(cases.GOLDEN, "lepl8_V.txt")).read()


@pytest.mark.parametrize("name", list(cases.CASES))
def test_port_matches_reference_outputs(name, golden):
	problem = cases.build_oracle_only(name)
	system = problem.system()

	assert np.array_equal(system.b, golden[f"{name}/b"])
	assert int(np.count_nonzero(system.val)) == int(golden[f"{name}/nnz"])
	assert system.bandwidth() == int(golden[f"{name}/bandwidth_natural"])

	if cases.CASES[name]:
		A = system.dense()
		assert hashlib.sha1(np.packbits(A != 0).tobytes()).hexdigest() == str(golden[f"{name}/pattern_sha1"])
		assert np.abs(A).sum() == float(golden[f"{name}/abs_sum"])

	perm, inv_perm = system.rcm()

	assert np.array_equal(perm.astype(np.int64), golden[f"{name}/perm"])
	assert system.bandwidth(perm) == int(golden[f"{name}/bandwidth_rcm"])

	x = system.band_solve(perm).reshape(-1, 2)

	assert np.array_equal(x, golden[f"{name}/effects"])  # bit for bit


@pytest.mark.parametrize("name", [n for n in cases.CASES if n not in cases.HEAVY])
def test_port_matches_live_reference(name, ref):
	case = cases.build(name, ref)
	case.sim.run()

	x = cases.oracle_problem(case).run()

	assert np.array_equal(x, case.instance.effects)
