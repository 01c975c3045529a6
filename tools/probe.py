"""GPU-box probe: stage timings of the plate workload at a few sizes (development aid, not a benchmark)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cases
from bfm_b200 import api, ext

lib = api.default_binding()
sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(400, 100), (1000, 250), (2000, 500)]

for nx, ny in sizes:
	t = time.time(); mesh = ext.plate(nx, ny, binding=lib); t_mesh = time.time() - t
	x = mesh.coords_array[:, 0]
	left = x == 0.0
	conds = [(0, 0.0, left), (1, 0.0, left)]
	case = cases._assemble_case("p", lib, mesh, 2, cases.STEEL, [cases.GRAVITY], conds)
	t = time.time(); job = ext.Job(case.sim); t_create = time.time() - t
	job.upload()
	for rep in range(2):
		t = time.time(); job.assemble(); job.solve(); wall = time.time() - t
	s = job.stats()
	spmv = job.spmv_ms(50)
	n = s["n_dofs"]; nnz = 4 * s["n_blocks"]
	canon = 12 * nnz + 20 * n + 4
	stored = s["n_slots"] * 36 + n // 2 * 32
	it_bytes = stored + 72 * n  # update_xr 48 B/DOF + update_p 24 B/DOF
	print(f"{nx}x{ny}: n={n} mesh {t_mesh:.2f}s create {t_create:.2f}s (plan {s['ms_plan']:.0f} ms) upload {s['ms_upload']:.2f} ms asm {s['ms_assemble']:.3f} ms bc {s['ms_bc']:.3f} ms "
	      f"solve {s["ms_solve"]:.1f} ms (setup {s["ms_solve_setup"]:.1f}, coarse {s["coarse_dim"]}) wall {wall*1e3:.1f} ms iters {s['cg_iterations']} res {s['cg_rel_residual']:.2e} true {s['cg_true_rel_residual']:.2e} restarts {s['cg_restarts']} "
	      f"us/iter {s['ms_solve']*1e3/max(s['cg_iterations'],1):.1f} spmv {spmv*1e3:.1f} us -> canonical {canon/spmv/1e6:.0f} GB/s stored {stored/spmv/1e6:.0f} GB/s; iter stored-bytes {it_bytes/(s['ms_solve']/max(s['cg_iterations'],1))/1e6:.0f} GB/s", flush=True)
	t = time.time(); case.sim.run(); print(f"   sim.run wall {time.time()-t:.3f}s", ext.last_stats(lib)["ms_download"], flush=True)
	del job, case, mesh
