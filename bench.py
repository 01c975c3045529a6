#!/usr/bin/env python
"""bench.py - assembly + solve throughput of the FEM hot path on the synthetic plate.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells NXxNY] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over the plate: element assembly + boundary conditions + the FP64
PCG solve to a relative residual of 1e-12.  Rank 0 prints ONE JSON line:

  value         DOF/s of assembly + solve with the mesh already resident in HBM (bfmx_job_assemble +
                bfmx_job_solve), timed with CUDA events on the library's stream, max over ranks
  e2e           the same metric through the reference-facing call, bfm_sim_run, with HOST buffers:
                host->device copies of coordinates / BC lists and the device->host copy of the
                displacements are inside the timed region
  roofline      the dominant kernel (the CG SpMV, k_spmv<kDot>): algorithmic bytes of the stored SELL-32
                2x2-block format per launch / its measured launch duration, against MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference libbfm (oracle/_ref, compiled from /root/reference by
                oracle/Makefile) timed on one host core on a bounded sample of the same workload

N > 1: the plate is row-partitioned over the ranks (strong scaling: the total mesh is fixed), with the
interface-DOF halo exchange and the CG dot products going over NVLink (bfm_b200/csrc/dist.cu).

`--impl reference` times the reference's own CPU implementation of the path (rank 0 only).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "assembly+solve throughput"
UNIT = "DOF/s"

DEFAULT_CELLS = "10000x2500"      # 10001 x 2501 nodes = 50.0 M DOF, the north star's plate: 6.3 GB of matrix, 50x the 126 MB L2
REFERENCE_SAMPLE = "160x40"       # 13 202 DOF: the largest SURVEY.md parity size the dense reference does in ~10 s


def parse_cells(text: str) -> tuple[int, int]:
	nx, ny = (int(v) for v in text.lower().split("x"))
	return nx, ny


# ---- clocks ------------------------------------------------------------------------------------


class ClockSampler:
	"""nvidia-smi clocks / throttle reasons of one GPU while the timed region runs (B200_PROFILING.md)"""

	FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

	def __init__(self, index: int):
		self.index = index
		self.proc = None
		self.lines: list[str] = []
		self.thread = None

	def start(self):
		try:
			self.proc = subprocess.Popen(
				["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
			)
		except OSError:
			self.proc = None
			return

		def pump():
			for line in self.proc.stdout:
				self.lines.append(line)

		self.thread = threading.Thread(target=pump, daemon=True)
		self.thread.start()

	def stop(self) -> dict:
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}

		self.proc.terminate()

		try:
			self.proc.wait(timeout=5)
		except subprocess.TimeoutExpired:
			self.proc.kill()

		self.thread.join(timeout=5)

		sm, sm_max, power, reasons = [], [], [], set()
		names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

		for line in self.lines:
			parts = [p.strip() for p in line.split(",")]

			if len(parts) < 7:
				continue

			try:
				sm.append(float(parts[0]))
				sm_max.append(float(parts[1]))
				power.append(float(parts[2]))
			except ValueError:
				continue

			for name, flag in zip(names, parts[3:7]):
				if flag.lower().startswith("active"):
					reasons.add(name)

		# "under load" = samples drawing more than half of the highest power seen
		busy = [s for s, p in zip(sm, power) if power and p >= 0.5 * max(power)] or sm

		return {
			"sm_mhz": statistics.median(busy) if busy else None,
			"sm_max_mhz": max(sm_max) if sm_max else None,
			"power_w_max": max(power) if power else None,
			"samples": len(sm),
			"reasons": sorted(reasons),
		}


# ---- the reference arm / CPU baseline ------------------------------------------------------------


def time_reference(cells: str, steps: int, warmup: int) -> dict:
	"""the UNMODIFIED reference libbfm (oracle/_ref/libbfm_ref.so), bfm_sim_run on one core"""

	from bfm_b200 import workloads
	from oracle import ref

	if not ref.available():
		raise RuntimeError("oracle/_ref/libbfm_ref.so is missing (python -c 'import __graft_entry__ as g; g.build()' builds it where /root/reference exists)")

	nx, ny = parse_cells(cells)
	case = workloads.plate_case(nx, ny, binding=ref.binding(), native_mesh=False)

	for _ in range(warmup):
		case.sim.run()

	t0 = time.perf_counter()

	for _ in range(steps):
		case.sim.run()

	seconds = (time.perf_counter() - t0) / steps

	return {
		"value": case.n_dofs / seconds,
		"unit": UNIT,
		"cores": 1,  # libbfm is single-threaded
		"kind": "reference",
		"sample": f"plate {nx}x{ny} cells = {case.n_dofs} DOF (dense reference caps at 46 340 DOF), bfm_sim_run x{steps}, {seconds:.3f} s each, host has {os.cpu_count()} logical cores",
		"seconds_per_run": seconds,
	}


def reference_arm(args, rank: int):
	if rank != 0:
		return

	nx, ny = parse_cells(args.cells)
	sx, sy = parse_cells(args.reference_sample)
	base = time_reference(args.reference_sample, args.steps, args.warmup)

	# the dense reference cannot hold the GPU arm's plate (8 n^2 bytes, int index: n <= 46 340), so its line
	# describes the plate it really ran; the GPU arm reports its own figure at this same size as `same_size`
	config = workload_config(sx, sy, 1)
	config["solver"] = "reference libbfm: dense assembly, RCM, band LU (unmodified, oracle/_ref)"
	config["partition"] = "none (single-threaded CPU code)"
	config["l2"] = "n/a (CPU)"
	config["reference_sample"] = base["sample"]
	config["gpu_arm_workload"] = workload_config(nx, ny, args.gpus)["workload"]

	print(json.dumps({
		"impl": "reference",
		"metric": METRIC,
		"value": base["value"],
		"unit": UNIT,
		"n_gpus": args.gpus,
		"steps": args.steps,
		"warmup": args.warmup,
		"ms_per_step": base["seconds_per_run"] * 1e3,
		"higher_is_better": True,
		"scaling": "strong",
		"vs_baseline": None,
		"dtype": "f64",
		"data": "synthetic",
		"config": config,
		"same_config": False,  # a cross-size ratio: the reference's DOF/s falls like 1/n, see `same_size` on the GPU arm's line
		"cpu_baseline": base,
		"e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}), flush=True)


def workload_config(nx: int, ny: int, gpus: int, stats: dict | None = None) -> dict:
	solver = "FP64 PCG to a relative residual of 1e-12, Jacobi scaling"

	if stats and stats.get("mg_levels", 0):
		smoothed = os.environ.get("BFM_MG_SMOOTH", "1") != "0"
		solver += f" + {'smoothed-aggregation multigrid V-cycle' if smoothed else 'aggregation multigrid W-cycle'} preconditioner ({stats['mg_levels']} levels, rigid-body modes per aggregate, damped-Jacobi smoothing fused into the SpMVs, dense last level of dimension {stats.get('coarse_dim', 0)})"
	elif stats and stats.get("coarse_dim", 0):
		solver += f" + one additive coarse level of rigid-body modes ({stats['coarse_dim'] // 3} aggregates, dense inverse)"

	return {
		"workload": f"synthetic structured triangulated plate {nx}x{ny} cells, {2 * (nx + 1) * (ny + 1)} DOF, plane stress, steel, gravity, left edge clamped (BASELINE.json configs[3])",
		"cells": f"{nx}x{ny}",
		"n_dofs": 2 * (nx + 1) * (ny + 1),
		"element": "P1 triangle, 3-point Gauss",
		"solver": solver,
		"partition": "none" if gpus == 1 else f"{gpus} contiguous node-row blocks, halo exchange + dot-product reductions over NVLink",
		"l2": "inputs larger than L2 (matrix 6.3 GB at the default size, 126 MB of L2); no flush needed",
	}



# ---- algorithmic bytes of one multigrid-preconditioned PCG iteration (mg.cuh header) ----


def mg_gammas(n_levels: int, smoothed: bool = True) -> list[int]:
	"""visits of the next level per cycle on the sparse levels 1 .. n_levels - 2 (mg.cuh: BFM_MG_GAMMA; default 1 -
	a V-cycle - with smoothed aggregation, 2 with the tentative prolongator)"""

	env = os.environ.get("BFM_MG_GAMMA", "")
	default = 1 if smoothed else 2
	out = []

	for l in range(1, n_levels - 1):
		g = default

		if env:
			parts = env.split(",")
			g = int(parts[min(l - 1, len(parts) - 1)] or default)

		out.append(g if 1 <= g <= 4 else default)

	return out


def mg_iteration_bytes(levels: list[dict], coarse_dim: int, smoothed: bool = True) -> dict:
	"""levels: [{n, n_slots, n_entries}] from bfmx_hier_info (E = entries of the level's prolongator: n with the
	tentative one, ~2.7 n with smoothed aggregation).  Stored-format bytes every kernel of one iteration must move:
	level 0 (2x2 blocks; S slots): k_spmv<kDot> 36 S + 32 n (FP64 operator), k_update_xr 96 n, k_spmv_mg<kPre>
	20 S + 32 n and k_spmv_mg<kPost> 20 S + 48 n (FP32 copy of the operator), k_mg_restrict 16 n + 32 E (t, P in FP32
	24 B + two indices), k_mg_prolong 32 n + 28 E (r, z, P 24 B + column), k_update_p 48 n;
	level l >= 1 (3x3 blocks, FP32 copy: 40 B per slot; gamma visits of the next level): (gamma + 1) fused SpMVs of
	40 S + 72 n, gamma restrictions of 24 n + 80 E and prolongations of 48 n + 76 E (+ 24 n when adding) and the next
	level's vector (24 B per node) twice; the dense last level 8 nc^2 per visit."""

	n0, s0, e0 = levels[0]["n"], levels[0]["n_slots"], levels[0]["n_entries"]
	n1 = levels[1]["n"] if len(levels) > 1 else 0
	per_level = [(36 + 20 + 20) * s0 + (32 + 96 + 32 + 48 + 16 + 32 + 48) * n0 + (32 + 28) * e0 + 2 * 24 * n1]
	gammas = mg_gammas(len(levels), smoothed)
	visits = 1

	for l in range(1, len(levels) - 1):
		n, sl, e, g = levels[l]["n"], levels[l]["n_slots"], levels[l]["n_entries"], gammas[l - 1]
		per_level.append(visits * ((g + 1) * (40 * sl + 72 * n) + g * ((24 + 48) * n + (80 + 76) * e) + (g - 1) * 24 * n + g * 2 * 24 * levels[l + 1]["n"]))
		visits *= g

	nc = (coarse_dim + 31) // 32 * 32
	per_level.append(visits * 8 * nc * nc)

	return {"total": sum(per_level), "per_level": per_level, "gammas": gammas, "dense_visits": visits, "smoothed_aggregation": smoothed}


# ---- parity where the driver can see it ---------------------------------------------------------------------


PARITY_CASES = ["plate_80x20", "plate_q4_24x6", "lepl8_all_kinds", "gear60"]
PARITY_TOL = 1e-9  # BASELINE.json north_star: displacements within 1e-9 (relative L2) of the reference's own solve


def parity_check(binding, rank: int, world: int) -> dict:
	"""bfm_sim_run on fixture cases (tests/cases.py) against the committed outputs of the UNMODIFIED reference
	(tests/golden/ref_outputs.npz) - on N > 1 GPUs through the partitioned path, collectively, every rank holding the
	complete field afterwards.  A miss aborts the run with a non-zero exit code: a throughput number for wrong
	displacements is worthless."""

	import numpy as np

	sys.path.insert(0, os.path.join(ROOT, "tests"))

	import cases
	from bfm_b200 import ext

	golden = cases.golden()
	errors, details = {}, {}

	for name in PARITY_CASES:
		case = cases.build(name, binding)
		case.sim.run()
		stats = ext.last_stats(binding)
		want = golden[f"{name}/effects"]
		errors[name] = float(np.linalg.norm(case.instance.effects - want) / np.linalg.norm(want))
		details[name] = {"iterations": stats["cg_iterations"], "mg_levels": stats["mg_levels"], "n_ranks": stats["n_ranks"], "converged": stats["cg_converged"]}

		if not (errors[name] <= PARITY_TOL and stats["cg_converged"] == 1 and stats["n_ranks"] == world):
			raise SystemExit(f"[rank {rank}] parity check failed on {name}: relative L2 error {errors[name]:.3e} (tolerance {PARITY_TOL}), stats {stats}")

	return {"tolerance": PARITY_TOL, "reference": "tests/golden/ref_outputs.npz (unmodified reference libbfm)", "rel_l2": errors, "runs": details}


# ---- configs[4]: 1024 batched small systems ---------------------------------------------------------


def batch_workload(args, binding):
	"""1024 variants of config 1 (8.lepl1110 + problem.txt, 670 DOF each; Young's modulus swept): the
	batched form of examples/benchmark.py's loop.  One assembly launch, one CTA per system."""

	import numpy as np

	from bfm_b200 import ext, workloads

	lib = binding.lib
	golden = os.path.join(ROOT, "tests", "golden")
	cases_ = workloads.lepl_sweep(os.path.join(golden, "meshes", "8.lepl1110"), os.path.join(golden, "problems", "problem.txt"), args.batch, binding)
	sims = [c.sim for c in cases_]

	job = ext.Job.batch(sims)
	job.upload()

	for _ in range(args.warmup):
		job.assemble()
		job.solve()

	sampler = ClockSampler(0)
	sampler.start()
	assert not lib.bfmx_device_sync()

	launches0 = lib.bfmx_kernel_launches()
	assert not lib.bfmx_timer_start(0)

	for _ in range(args.steps):
		job.assemble()
		job.solve()

	ms = lib.bfmx_timer_stop(0) / args.steps
	launches = lib.bfmx_kernel_launches() - launches0
	clocks = sampler.stop()
	s = job.stats()
	status = job.batch_status()
	n_dofs = s["n_dofs"]

	# the PCG kernel streams each system's scaled matrix once per iteration (L2-resident at this size)
	iters_total = sum(st["iterations"] for st in status)
	sys_bytes = s["n_slots"] * 36 / len(status)
	solve_bytes = iters_total * sys_bytes

	ext.sim_run_batch(sims)  # warm-up of the end-to-end call
	t0 = time.perf_counter()
	h2d = d2h = 0

	for _ in range(args.steps):
		ext.sim_run_batch(sims)
		st = ext.last_stats(binding)
		h2d += st["h2d_bytes"]
		d2h += st["d2h_bytes"]

	e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
	check = float(np.abs(workloads.effects_view(cases_[0].instance)).max())

	cpu = None

	if not args.no_cpu_baseline:
		from oracle import ref

		ref_cases = workloads.lepl_sweep(os.path.join(golden, "meshes", "8.lepl1110"), os.path.join(golden, "problems", "problem.txt"), 16, ref.binding())
		t0 = time.perf_counter()

		for c in ref_cases:
			c.sim.run()

		sec = (time.perf_counter() - t0) / len(ref_cases)
		cpu = {"value": 670 / sec, "unit": UNIT, "cores": 1, "kind": "reference", "sample": f"16 of the {args.batch} systems, bfm_sim_run each, {sec * 1e3:.2f} ms per system (what examples/benchmark.py times)"}

	print(json.dumps({
		"metric": METRIC, "value": n_dofs / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": f"{len(status)} batched small systems (8.lepl1110 + problem.txt with E swept, 670 DOF each), one CTA per system (BASELINE.json configs[4])", "systems": len(status), "n_dofs": n_dofs},
		"assembly_ms": s["ms_assemble"] + s["ms_bc"], "solve_ms": s["ms_solve"],
		"cg_iterations_max": s["cg_iterations"], "cg_iterations_total": iters_total, "systems_per_s": len(status) / (ms * 1e-3),
		"roofline": {"kernel": "k_pcg_cta<256> (whole PCG of one system per CTA)", "bound": "l2", "achieved": solve_bytes / (s["ms_solve"] * 1e-3) / 1e9, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
			"note": "matrix bytes streamed per second from L2/L1 (the batch's matrices total %.0f MB, inside the 126 MB L2); no HBM roofline applies" % (s["n_slots"] * 36 / 1e6)},
		"cpu_baseline": cpu,
		"e2e": {"value": n_dofs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps, "ms_per_step": e2e_ms, "call": "bfmx_sim_run_batch (plan build + upload + assemble + solve + download, every step)", "max_abs_displacement": check},
		"gpu_launches": launches, "clocks": clocks,
	}), flush=True)


# ---- our arm ----------------------------------------------------------------------------------------


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=3)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
	ap.add_argument("--cells", default=DEFAULT_CELLS, help="plate size NXxNY (cells)")
	ap.add_argument("--reference-sample", default=REFERENCE_SAMPLE, help="plate size the CPU reference is timed on")
	ap.add_argument("--workload", default="plate", choices=["plate", "batch"], help="plate: BASELINE.json configs[3] (the headline); batch: configs[4], 1024 small systems, one CTA each")
	ap.add_argument("--batch", type=int, default=1024, help="systems in the batch workload")
	ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
	ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
	ap.add_argument("--no-parity-check", action="store_true", help="skip the fixture parity check that precedes the timing")
	args = ap.parse_args()

	rank = int(os.environ.get("RANK", "0"))
	world = int(os.environ.get("WORLD_SIZE", "1"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))

	# host-side symbolic work (sparsity plan, partition, aggregates) is OpenMP code in libbfm; torchrun pins
	# OMP_NUM_THREADS to 1 unless told otherwise, so share the host cores between the ranks instead
	os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(world, 1)))

	if args.impl == "reference":
		reference_arm(args, rank)
		return

	if world != args.gpus:
		raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")

	import numpy as np

	from bfm_b200 import api, ext, workloads

	binding = api.default_binding()  # raises when libbfm.so is not built: there is no fallback

	if not ext.device_available(binding):
		raise SystemExit("no CUDA device: " + binding.lib.bfmx_device_error().decode())

	lib = binding.lib
	dist = None

	if world > 1:
		import torch
		import torch.distributed as dist

		# NCCL announces its version on stdout when a communicator comes up; stdout is reserved for the one
		# JSON line, so it points at stderr until the communicators exist
		sys.stdout.flush()
		saved_stdout = os.dup(1)
		os.dup2(2, 1)

		try:
			torch.cuda.set_device(local_rank)
			dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
			dist.barrier()
			ext.dist_init(binding, dist, local_rank)  # hands the library its own communicator (bfmx_dist_init)

			if rank == 0 and binding.lib.bfmx_dist_peer_memory_status():
				sys.stderr.write("peer memory off: " + binding.lib.bfmx_dist_peer_memory_status().decode() + "\n")
			torch.cuda.synchronize()
		finally:
			sys.stdout.flush()
			os.dup2(saved_stdout, 1)
			os.close(saved_stdout)

	def barrier():
		if dist is not None:
			dist.barrier()

		assert not lib.bfmx_device_sync()

	def max_over_ranks(value: float) -> float:
		if dist is None:
			return value

		import torch

		t = torch.tensor([value], dtype=torch.float64, device="cuda")
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		return float(t.item())

	if args.workload == "batch":
		if world != 1:
			raise SystemExit("the batch workload is embarrassingly parallel: run one process per GPU on its own share")

		batch_workload(args, binding)
		return

	parity = None if args.no_parity_check else parity_check(binding, rank, world)

	nx, ny = parse_cells(args.cells)
	case = workloads.plate_case(nx, ny, binding=binding)
	n_dofs = case.n_dofs

	t_create = time.perf_counter()
	job = ext.Job(case.sim)  # symbolic phase, once per mesh: sparsity plan, row partition, aggregates (host) + their upload
	symbolic_ms = (time.perf_counter() - t_create) * 1e3
	job.upload()  # inputs resident in HBM from here on

	def step():
		job.assemble()
		job.solve()

	for _ in range(args.warmup):
		step()

	sampler = ClockSampler(local_rank)

	if rank == 0:
		sampler.start()

	barrier()

	launches0 = lib.bfmx_kernel_launches()
	per_step = []

	assert not lib.bfmx_timer_start(0)

	for _ in range(args.steps):
		step()
		per_step.append(job.stats())

	ms_total = lib.bfmx_timer_stop(0)  # synchronises
	launches = lib.bfmx_kernel_launches() - launches0

	barrier()

	ms_total = max_over_ranks(ms_total)
	clocks = sampler.stop() if rank == 0 else None

	ms_per_step = ms_total / args.steps
	value = n_dofs / (ms_per_step * 1e-3)

	s = per_step[-1]
	iters = s["cg_iterations"]
	ms_asm = statistics.mean(p["ms_assemble"] + p["ms_bc"] for p in per_step)
	ms_solve = statistics.mean(p["ms_solve"] for p in per_step)

	# ---- roofline of the dominant kernel: the CG SpMV, timed live with CUDA events on the library stream

	n_own = s["n_dofs_owned"] if "n_dofs_owned" in s else s["n_dofs"]
	spmv_ms = max_over_ranks(job.spmv_ms(50))

	stored_bytes = s["n_slots"] * 36 + (n_own // 2) * 32             # 2x2 blocks (32 B) + block column (4 B); p gathered once, q written once
	canonical_bytes = 12 * 4 * s["n_blocks"] + 20 * n_own + 4        # scalar CSR, 12 B/nnz + 20 B/row
	iter_bytes = stored_bytes + 72 * n_own                           # + update_xr (6 x 8 B/DOF) + update_p (3 x 8 B/DOF)

	mg_model = None

	if s.get("mg_levels", 0) and world == 1:
		import ctypes

		info = ext.HierInfo()
		assert not lib.bfmx_hier_info(case.mesh.c_mesh, ctypes.byref(info))
		levels = [{"n": info.n_nodes[l], "n_slots": info.n_slots[l], "n_entries": info.n_entries[l]} for l in range(info.n_levels)]
		mg_model = mg_iteration_bytes(levels, s["coarse_dim"], bool(info.smoothed))
		mg_model["levels"] = levels
		iter_bytes = mg_model["total"]

	elif s.get("coarse_dim", 0):
		# two-level preconditioner: restriction (index 4 B + r 16 B + W row 16 B per node), E^-1 g (dense, nc x nc doubles),
		# prolongation fused in the p update (+ W row 16 B + aggregate id 4 B per node)
		nc = (s["coarse_dim"] + 31) // 32 * 32
		apply_rows = nc // world if s.get("uses_peer_memory") else nc  # over peer memory each rank applies its own rows of E^-1
		iter_bytes += (36 + 20) * (n_own // 2) + 8 * nc * apply_rows

	peaks = {}
	peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")

	if os.path.exists(peaks_path):
		peaks = json.load(open(peaks_path))

	peak = float(peaks.get("hbm_gbs", 6650.0))
	peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

	# DRAM traffic of the same kernel on the same workload from the committed ncu --set full capture (one GPU)
	traffic = None
	traffic_path = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")

	if world == 1 and os.path.exists(traffic_path):
		traffic = (json.load(open(traffic_path)).get(args.cells) or {}).get("dram_bytes_per_launch")

	achieved = stored_bytes / (spmv_ms * 1e-3) / 1e9
	ms_setup = statistics.mean(p.get("ms_solve_setup", 0.0) for p in per_step)
	us_iter = (ms_solve - ms_setup) * 1e3 / max(iters, 1)  # the iteration loop alone: scaling + coarse set-up are reported apart

	roofline = {
		"kernel": "k_spmv<kDot> (q = A p over SELL-32 2x2 node blocks, fused p.q)",
		"bound": "hbm",
		"achieved": achieved,
		"peak": peak,
		"unit": "GB/s",
		"frac": achieved / peak,
		"frac_of_nominal_8TBs": achieved / 8000.0,
		"peak_source": peak_src,
		"traffic": traffic,  # dram__bytes_read + write per launch, ncu --set full (profiles/r1_ncu_traffic.json); None when not captured for this workload
		"bytes_per_launch": stored_bytes,
		"us_per_launch": spmv_ms * 1e3,
		"canonical_csr_bytes_per_launch": canonical_bytes,
		"canonical_csr_gbs": canonical_bytes / (spmv_ms * 1e-3) / 1e9,
		"cg_iteration": {
			"us": us_iter,
			"bytes": iter_bytes,
			"achieved": iter_bytes / (us_iter * 1e-6) / 1e9,
			"frac": iter_bytes / (us_iter * 1e-6) / 1e9 / peak,
			"note": "whole PCG iteration (CG SpMV + vector kernels + the preconditioner's fused SpMVs, restrictions, prolongations and coarser levels) over the timed region: (solve ms - set-up ms) / iterations",
			"model": mg_model,
		},
	}

	# assembly: one launch of k_assemble + the boundary-condition kernels.  Algorithmic bytes: connectivity
	# 4 B x kind x elements, coordinates 16 B per node, contributor map 4 B per contribution (9 per P1 element) + 4 B per
	# slot of list bounds, 32 B written per 2x2 block slot, 16 B of load vector per node
	n_nodes_own = n_own // 2
	n_elems = 2 * nx * ny // world
	asm_bytes = 4 * 3 * n_elems + 16 * n_nodes_own + 4 * 9 * n_elems + 4 * s["n_slots"] + 32 * s["n_slots"] + 16 * n_nodes_own
	ms_asm_kernel = statistics.mean(p["ms_assemble"] for p in per_step)

	roofline_assembly = {
		"kernel": "k_assemble<3,0,0> (one lane per 2x2 block walks its contributor list in the reference's order)",
		"bound": "hbm",
		"achieved": asm_bytes / (ms_asm_kernel * 1e-3) / 1e9,
		"peak": peak,
		"unit": "GB/s",
		"frac": asm_bytes / (ms_asm_kernel * 1e-3) / 1e9 / peak,
		"bytes_per_launch": asm_bytes,
		"us_per_launch": ms_asm_kernel * 1e3,
		"note": "latency- and FP64-bound rather than HBM-bound: every block recomputes its elements' geometry (bit-for-bit reference arithmetic, -fmad=false, true divisions); 1 % of a step",
	}

	# ---- end to end: bfm_sim_run on host buffers (the call pybfm makes), copies inside the timed region

	e2e = None

	if not args.no_e2e:
		# warm-up.  The library keeps what depends on the mesh alone - sparsity plan, row partition, aggregates -
		# across calls on the same mesh object (examples/benchmark.py calls sim.run() ten times on one mesh);
		# numbers never are: every call assembles, sets up the coarse operator and solves from scratch.
		t_first = time.perf_counter()
		case.sim.run()
		first_call_ms = (time.perf_counter() - t_first) * 1e3
		barrier()

		t0 = time.perf_counter()
		assert not lib.bfmx_timer_start(1)
		h2d = d2h = 0
		stages = {"ms_plan": 0.0, "ms_upload": 0.0, "ms_assemble": 0.0, "ms_bc": 0.0, "ms_solve": 0.0, "ms_download": 0.0}

		for _ in range(args.steps):
			case.sim.run()
			st = ext.last_stats(binding)
			h2d += st["h2d_bytes"]
			d2h += st["d2h_bytes"]

			for key in stages:
				stages[key] += st[key] / args.steps

		# the displacements are on the host now (the copy into instance->effects is part of every bfm_sim_run); read the
		# tip like a caller would.  The full max-norm below is a check on 400 MB of host memory (~120 ms in numpy at
		# 50 M DOF), not part of the path: it runs after the clock has stopped
		tip = float(workloads.effects_view(case.instance)[-1])

		ms_e2e = lib.bfmx_timer_stop(1)
		wall = (time.perf_counter() - t0) * 1e3
		barrier()

		view = workloads.effects_view(case.instance)
		checksum = float(max(view.max(), -view.min()))
		assert np.isfinite(tip) and abs(tip) <= checksum

		ms_e2e = max_over_ranks(max(ms_e2e, wall)) / args.steps

		e2e = {
			"value": n_dofs / (ms_e2e * 1e-3),
			"unit": UNIT,
			"h2d_bytes_per_step": h2d // args.steps,
			"d2h_bytes_per_step": d2h // args.steps,
			"ms_per_step": ms_e2e,
			"call": "bfm_sim_run (job create with cached symbolic plan + upload + assemble + solve + download into instance->effects)",
			"first_call_ms": first_call_ms,
			"symbolic_setup_ms_once_per_mesh": symbolic_ms,
			"max_abs_displacement": checksum,
			"stages_ms": stages | {"host_side_rest": ms_e2e - sum(stages.values())},  # rank 0's stages; the rest is host work of job creation (cache look-ups, BC work lists)
		}

	cpu = None
	same_size = None

	if rank == 0 and world == 1 and not args.no_cpu_baseline:
		cpu = time_reference(args.reference_sample, 1, 0)

		# the GPU arm on the very plate the reference arm runs (the dense reference cannot go beyond 46 340 DOF):
		# the same-work ratio next to the cross-size headline, and parity at the reference's own size
		sx, sy = parse_cells(args.reference_sample)
		small = workloads.plate_case(sx, sy, binding=binding)
		small.sim.run()
		t0 = time.perf_counter()
		reps = 20

		for _ in range(reps):
			small.sim.run()

		seconds = (time.perf_counter() - t0) / reps
		same_size = {
			"cells": args.reference_sample, "n_dofs": small.n_dofs, "value": small.n_dofs / seconds, "unit": UNIT, "ms_per_step": seconds * 1e3,
			"call": "bfm_sim_run on host buffers (one CTA solves a mesh this small)", "reference_value": cpu["value"], "ratio": small.n_dofs / seconds / cpu["value"],
		}

		golden_path = os.path.join(ROOT, "tests", "golden", "ref_outputs.npz")
		key = f"plate_{args.reference_sample}/effects"

		if os.path.exists(golden_path):
			golden = np.load(golden_path)

			if key in golden:
				got = workloads.effects_view(small.instance).reshape(-1, 2)
				same_size["rel_l2_vs_reference"] = float(np.linalg.norm(got - golden[key]) / np.linalg.norm(golden[key]))

	if rank == 0:
		print(json.dumps({
			"metric": METRIC,
			"value": value,
			"unit": UNIT,
			"n_gpus": world,
			"steps": args.steps,
			"warmup": args.warmup,
			"ms_per_step": ms_per_step,
			"higher_is_better": True,
			"scaling": "strong",
			"vs_baseline": None,  # the reference publishes no number for this metric
			"dtype": "f64",
			"data": "synthetic",
			"config": workload_config(nx, ny, world, s),
			"assembly_ms": ms_asm,
			"solve_ms": ms_solve,
			"symbolic_setup_ms_once_per_mesh": symbolic_ms,
		"cg_iterations": iters,
		"coarse_dim": s.get("coarse_dim", 0),
		"exchange": ("NVLink peer memory (CUDA IPC mailboxes), posted from inside the kernels" if s.get("uses_peer_memory") else "NCCL") if world > 1 else "none",
		"solve_setup_ms": s.get("ms_solve_setup", 0.0),
			"mg_levels": s.get("mg_levels", 0),
			"cg_rel_residual": s["cg_rel_residual"],            # recursive residual CG stops on (<= 1e-12)
			"cg_true_rel_residual": s["cg_true_rel_residual"],  # recomputed b - A x: floors at ~eps * cond(A) in FP64 for any solver
			"cg_backward_error": s["cg_backward_error"],        # ||b - A x|| / (||x|| + ||b||), scaled norm: what a backward-stable solve guarantees
			"cg_restarts": s["cg_restarts"],
			"roofline": roofline,
			"roofline_assembly": roofline_assembly,
			"cpu_baseline": cpu,
			"same_size": same_size,
			"parity_check": parity,
			"e2e": e2e,
			"gpu_launches": launches,
			"clocks": clocks,
		}), flush=True)

	if dist is not None:
		ext.dist_finalize(binding)
		dist.destroy_process_group()


if __name__ == "__main__":
	main()
