"""The symbolic phase of a plate on the device, for ncu: sparsity plan (symbolic.cu) + edge derivation.

    BFM_QUIET=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/profile_symbolic.py [NXxNY]
"""

import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bfm_b200 import api, ext  # noqa: E402

cells = sys.argv[1] if len(sys.argv) > 1 else "10000x2500"
nx, ny = (int(v) for v in cells.split("x"))

binding = api.default_binding()
assert ext.device_available(binding), binding.lib.bfmx_device_error()

mesh = ext.plate(nx, ny, binding=binding)
sizes = [C.c_size_t() for _ in range(4)]

t0 = time.perf_counter()
assert not binding.lib.bfmx_mesh_pattern_sizes(C.byref(mesh.c_mesh), *[C.byref(s) for s in sizes])
t1 = time.perf_counter()

os.environ["BFM_EDGES"] = "device"
assert not binding.lib.bfmx_mesh_compute_edges(C.byref(mesh.c_mesh))
t2 = time.perf_counter()

print(f"{cells}: {mesh.c_mesh.n_nodes} nodes, {sizes[1].value} slots, {sizes[3].value} contributions: plan {1e3 * (t1 - t0):.1f} ms; {mesh.c_mesh.n_edges} edges: {1e3 * (t2 - t1):.1f} ms (host wall clock, copies included)")
