"""Worker of tests/test_gpu_dist.py - one process per GPU (torchrun, NCCL): the partitioned hot path
against the committed reference outputs and the oracle.

For every case: bfm_sim_run as a collective call -> every rank holds the complete displacement field,
within 1e-9 (relative L2) of the reference's direct solve; the owned rows of the assembled right-hand
side equal the oracle's bit for bit (the local numbering is monotone, so the reference's accumulation
order is kept - partition.c)."""

import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

for p in (ROOT, os.path.join(ROOT, "tests")):
	sys.path.insert(0, p)

import cases  # noqa: E402
from bfm_b200 import api, ext  # noqa: E402


def main():
	names, out_path = sys.argv[1].split(","), sys.argv[2]
	local_rank = int(os.environ["LOCAL_RANK"])

	torch.cuda.set_device(local_rank)
	dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
	rank, world = dist.get_rank(), dist.get_world_size()

	lib = api.default_binding()
	assert ext.device_available(lib), lib.lib.bfmx_device_error()
	ext.dist_init(lib, dist, local_rank)

	assert lib.lib.bfmx_dist_world() == world and lib.lib.bfmx_dist_rank() == rank

	golden = cases.golden()
	report = {}

	for name in names:
		case = cases.build(name, lib)

		try:
			case.sim.run()
		except AssertionError:
			sys.stderr.write(f"[rank {rank}] bfm_sim_run failed on case {name}: {ext.last_stats(lib)}\n")
			raise

		stats = ext.last_stats(lib)

		if rank == 0:
			sys.stderr.write(f"case {name}: {stats['cg_iterations']} iterations, {stats['mg_levels']} multigrid levels, coarse {stats['coarse_dim']}, refinements {stats['cg_restarts']}, peer memory {stats['uses_peer_memory']}\n")
		u = case.instance.effects.copy()

		want = golden[f"{name}/effects"]
		err = float(np.linalg.norm(u - want) / np.linalg.norm(want))

		# every rank must hold the same bits
		mine = torch.from_numpy(u.reshape(-1).copy()).cuda()
		ref0 = mine.clone()
		dist.broadcast(ref0, 0)
		same = bool(torch.equal(mine, ref0))

		# staged: owned rows of b, bit for bit against the oracle (BFM_RENUMBER=1: the job partitions its internally
		# renumbered copy of the mesh, whose local rows are not the caller's - only the displacements are compared)
		job = ext.Job(case.sim)
		job.upload()
		job.assemble()

		if os.environ.get("BFM_RENUMBER") == "1":
			assert ext.internal_numbering(case.mesh) is not None
			b_equal = True

		else:
			part = ext.partition(case.mesh, rank, world)
			n_local = int(part["n_local_nodes"])
			b_local = np.zeros(2 * n_local)
			assert not lib.lib.bfmx_job_read(job.handle, b_local.ctypes.data_as(C_DOUBLE_P), None)
			oracle = cases.oracle_problem(case).system()
			own = slice(2 * int(part["own_begin"]), 2 * int(part["own_end"]))
			rows = slice(2 * int(part["first_node"]), 2 * int(part["end_node"]))
			b_equal = bool(np.array_equal(b_local[own], oracle.b[rows]))

		job.solve()
		job.download()
		staged_equal = bool(np.array_equal(case.instance.effects, u))

		report[name] = {
			"err": err, "same_on_all_ranks": same, "b_bitwise": b_equal, "staged_equals_run": staged_equal,
			"converged": stats["cg_converged"], "rel_residual": stats["cg_rel_residual"], "iterations": stats["cg_iterations"],
			"n_ranks": stats["n_ranks"], "peer_memory": stats["uses_peer_memory"], "coarse": stats["coarse_dim"], "owned": stats["n_dofs_owned"], "halo_bytes": stats["halo_bytes_per_exchange"],
		}

		del job, case

	gathered = [None] * world
	dist.all_gather_object(gathered, report)

	if rank == 0:
		with open(out_path, "w") as f:
			json.dump(gathered, f)

	ext.dist_finalize(lib)
	dist.destroy_process_group()


import ctypes  # noqa: E402

C_DOUBLE_P = ctypes.POINTER(ctypes.c_double)

if __name__ == "__main__":
	main()
