"""Short, fixed-length run of the hot path for ncu: assembly + BCs + `iters` PCG iterations on a plate.

    BFM_QUIET=1 ncu ... python tools/profile_target.py [NXxNY] [iters]

The CG loop is cut at `iters` iterations (BFM_CG_MAXIT), so the "not converged" return of the solve stage
is expected and ignored here; everything else is the product path of bench.py.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

cells = sys.argv[1] if len(sys.argv) > 1 else "2000x500"
iters = sys.argv[2] if len(sys.argv) > 2 else "100"

os.environ["BFM_CG_MAXIT"] = iters
os.environ["BFM_CG_CHUNK"] = iters
os.environ.setdefault("BFM_QUIET", "1")

from bfm_b200 import api, ext, workloads  # noqa: E402

binding = api.default_binding()
assert ext.device_available(binding), binding.lib.bfmx_device_error()

nx, ny = (int(v) for v in cells.split("x"))
case = workloads.plate_case(nx, ny, binding=binding)
job = ext.Job(case.sim)
job.upload()

for _ in range(2):
	job.assemble()
	binding.lib.bfmx_job_solve(job.handle)  # -1 = iteration limit, by construction

s = job.stats()
print(f"{cells}: {s['n_dofs']} DOF, {s['cg_iterations']} iterations, {s['kernel_launches']} launches, solve {s['ms_solve']:.2f} ms")
