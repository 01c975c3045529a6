"""Worker of tests/test_partition.py::test_two_rank_gloo_pcg - one process per rank over gloo (CPU).

Each rank asks the library for ITS partition of the mesh (bfmx_partition_*), keeps only its owned rows of
the oracle's system, and runs the emulated partitioned PCG with real inter-process halo exchange
(dist.send / dist.recv) and dot products (dist.all_reduce).  Rank 0 gathers the owned blocks and writes
the relative L2 error against the committed reference displacements."""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

for p in (ROOT, os.path.join(ROOT, "tests")):
	sys.path.insert(0, p)

import cases  # noqa: E402
import dist_emulation as emu  # noqa: E402
from bfm_b200 import api, ext  # noqa: E402


def main():
	name, out_path = sys.argv[1], sys.argv[2]

	dist.init_process_group("gloo")
	rank, world = dist.get_rank(), dist.get_world_size()

	lib = api.default_binding()
	case = cases.build(name, lib)
	oracle = cases.oracle_problem(case).system()

	view = emu.RankView(ext.partition(case.mesh, rank, world), oracle.scipy(), oracle.b.copy())

	def exchange(vectors):
		v = vectors[0]
		p = view.part
		reqs, recv = [], []

		for i, s in enumerate(p["nbr"]):
			reqs.append(dist.isend(torch.from_numpy(view.pack(v, i)), int(s)))
			buf = torch.zeros(2 * int(p["recv_count"][i]), dtype=torch.float64)
			reqs.append(dist.irecv(buf, int(s)))
			recv.append(buf)

		for req in reqs:
			req.wait()

		for i, buf in enumerate(recv):
			view.unpack(v, i, buf.numpy())

	def allsum(parts):
		t = torch.tensor([sum(parts)], dtype=torch.float64)
		dist.all_reduce(t)
		return float(t.item())

	xs, iters = emu.pcg([view], exchange, allsum)

	blocks = [None] * world
	dist.all_gather_object(blocks, xs[0])

	if rank == 0:
		x = np.concatenate(blocks)
		want = cases.golden()[f"{name}/effects"].reshape(-1)
		err = np.linalg.norm(x - want) / np.linalg.norm(want)

		with open(out_path, "w") as f:
			f.write(f"{err:.3e} {iters}\n")

	dist.destroy_process_group()


if __name__ == "__main__":
	main()
