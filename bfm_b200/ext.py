"""ctypes prototypes of the bfmx_* additions (include/bfm_b200.h) + small Python conveniences."""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as abi

_P = C.POINTER
_int = C.c_int
c_int32_p = _P(C.c_int32)
c_uint32_p = _P(C.c_uint32)


class Stats(C.Structure):  # bfmx_stats_t
	_fields_ = [
		("n_dofs", C.c_size_t),
		("n_blocks", C.c_size_t),
		("n_slots", C.c_size_t),
		("cg_iterations", _int),
		("cg_restarts", _int),
		("cg_converged", _int),
		("cg_rel_residual", C.c_double),
		("cg_true_rel_residual", C.c_double),
		("ms_plan", C.c_float),
		("ms_upload", C.c_float),
		("ms_assemble", C.c_float),
		("ms_bc", C.c_float),
		("ms_solve", C.c_float),
		("ms_download", C.c_float),
		("kernel_launches", C.c_size_t),
		("h2d_bytes", C.c_size_t),
		("d2h_bytes", C.c_size_t),
		("cg_backward_error", C.c_double),
		("n_dofs_owned", C.c_size_t),
		("n_ranks", C.c_size_t),
		("halo_bytes_per_exchange", C.c_size_t),
		("coarse_dim", C.c_size_t),
		("ms_solve_setup", C.c_float),
		("uses_peer_memory", _int),
		("mg_levels", _int),
	]

	def as_dict(self) -> dict:
		return {name: getattr(self, name) for name, _ in self._fields_}


class HierInfo(C.Structure):  # bfmx_hier_info_t
	_fields_ = [("n_levels", C.c_int32), ("n_nodes", C.c_int32 * 12), ("n_slots", C.c_int64 * 12), ("n_entries", C.c_int64 * 12), ("smoothed", C.c_int32), ("pad_", C.c_int32)]


class PartitionInfo(C.Structure):  # bfmx_partition_info_t
	_fields_ = [(name, C.c_size_t) for name in ("first_node", "end_node", "n_local_nodes", "own_begin", "own_end", "n_local_elems", "n_neighbours", "n_send")]


class BatchStatus(C.Structure):  # bfmx_batch_status_t
	_fields_ = [("iterations", _int), ("converged", _int), ("rel_residual", C.c_double), ("true_rel_residual", C.c_double), ("backward_error", C.c_double)]

	def as_dict(self) -> dict:
		return {name: getattr(self, name) for name, _ in self._fields_}


DIST_ID_BYTES = 128

PROTOTYPES = {
	"bfmx_dist_unique_id": (_int, [C.c_void_p]),
	"bfmx_dist_init": (_int, [_int, _int, C.c_void_p]),
	"bfmx_dist_finalize": (_int, []),
	"bfmx_dist_rank": (_int, []),
	"bfmx_dist_world": (_int, []),
	"bfmx_dist_peer_memory_status": (C.c_char_p, []),
	"bfmx_coarse_plan": (_int, [_P(abi.Mesh), _int, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
	"bfmx_hier_info": (_int, [_P(abi.Mesh), _P(HierInfo)]),
	"bfmx_hier_level": (_int, [_P(abi.Mesh), _int, c_int32_p, C.POINTER(C.c_float), c_int32_p, c_int32_p]),
	"bfmx_partition_sizes": (_int, [_P(abi.Mesh), _int, _int, _P(PartitionInfo)]),
	"bfmx_partition_copy": (_int, [_P(abi.Mesh), _int, _int, abi.c_size_t_p, abi.c_size_t_p, abi.c_size_t_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
	"bfmx_device_available": (_int, []),
	"bfmx_device_error": (C.c_char_p, []),
	"bfmx_device_sync": (_int, []),
	"bfmx_timer_start": (_int, [_int]),
	"bfmx_timer_stop": (C.c_float, [_int]),
	"bfmx_kernel_launches": (C.c_size_t, []),
	"bfmx_device_sm_count": (_int, []),
	"bfmx_last_stats": (_int, [_P(Stats)]),
	"bfmx_job_create": (_int, [_P(C.c_void_p), _P(abi.Sim), C.c_size_t]),
	"bfmx_job_upload": (_int, [C.c_void_p]),
	"bfmx_job_assemble": (_int, [C.c_void_p]),
	"bfmx_job_solve": (_int, [C.c_void_p]),
	"bfmx_job_download": (_int, [C.c_void_p]),
	"bfmx_job_stats": (_int, [C.c_void_p, _P(Stats)]),
	"bfmx_job_spmv_time": (_int, [C.c_void_p, _int, _P(C.c_float)]),
	"bfmx_job_read": (_int, [C.c_void_p, abi.c_double_p, abi.c_double_p]),
	"bfmx_job_destroy": (_int, [C.c_void_p]),
	"bfmx_batch_max_nodes": (_int, []),
	"bfmx_job_create_batch": (_int, [_P(C.c_void_p), _P(_P(abi.Sim)), C.c_size_t]),
	"bfmx_job_batch_size": (_int, [C.c_void_p]),
	"bfmx_job_batch_status": (_int, [C.c_void_p, C.c_size_t, _P(BatchStatus)]),
	"bfmx_sim_run_batch": (_int, [_P(_P(abi.Sim)), C.c_size_t]),
	"bfmx_mesh_compute_edges": (_int, [_P(abi.Mesh)]),
	"bfmx_mesh_internal_numbering": (_int, [_P(abi.Mesh), _P(C.c_int32)]),
	"bfmx_mesh_plate": (_int, [_P(abi.Mesh), _P(abi.State), C.c_size_t, C.c_size_t, C.c_double, C.c_double, _int, C.c_bool]),
	"bfmx_mesh_pattern_sizes": (_int, [_P(abi.Mesh), _P(C.c_size_t), _P(C.c_size_t), _P(C.c_size_t), _P(C.c_size_t)]),
	"bfmx_mesh_pattern_copy": (_int, [_P(abi.Mesh), c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_uint32_p]),
	"bfmx_matrix_csr_create": (_int, [_P(abi.Matrix), _P(abi.State), C.c_size_t, abi.c_size_t_p, abi.c_size_t_p, abi.c_double_p]),
	"bfmx_matrix_csr_export": (_int, [_P(abi.Matrix), _P(C.c_size_t), abi.c_size_t_p, abi.c_size_t_p, abi.c_double_p]),
}


def dist_init(binding, dist, local_rank: int | None = None):
	"""give the library its own NCCL communicator over the ranks of an initialised torch.distributed
	process group (used only to ship the 128-byte unique id; the data path never touches torch)"""

	rank, world = dist.get_rank(), dist.get_world_size()
	ident = C.create_string_buffer(DIST_ID_BYTES)

	if rank == 0:
		assert not binding.lib.bfmx_dist_unique_id(ident), binding.lib.bfmx_device_error()

	box = [ident.raw]
	dist.broadcast_object_list(box, src=0)
	ident = C.create_string_buffer(box[0], DIST_ID_BYTES)

	assert not binding.lib.bfmx_dist_init(rank, world, ident), binding.lib.bfmx_device_error()


def dist_finalize(binding):
	binding.lib.bfmx_dist_finalize()


def partition(mesh, rank: int, world: int) -> dict:
	"""the row partition rank `rank` of `world` would get for `mesh` (host-only, see partition.c)"""

	lib = mesh.binding.lib
	info = PartitionInfo()
	assert not lib.bfmx_partition_sizes(C.byref(mesh.c_mesh), rank, world, C.byref(info))

	kind = mesh.c_mesh.kind
	out = {
		"l2g": np.zeros(info.n_local_nodes, np.uint64),
		"elems": np.zeros((info.n_local_elems, kind), np.uint64),
		"elem_l2g": np.zeros(info.n_local_elems, np.uint64),
		"nbr": np.zeros(info.n_neighbours, np.int32),
		"recv_begin": np.zeros(info.n_neighbours, np.int32),
		"recv_count": np.zeros(info.n_neighbours, np.int32),
		"send_ptr": np.zeros(info.n_neighbours + 1, np.int32),
		"send_idx": np.zeros(max(info.n_send, 1), np.int32),
	}

	assert not lib.bfmx_partition_copy(
		C.byref(mesh.c_mesh), rank, world,
		out["l2g"].ctypes.data_as(abi.c_size_t_p), out["elems"].ctypes.data_as(abi.c_size_t_p), out["elem_l2g"].ctypes.data_as(abi.c_size_t_p),
		*[out[k].ctypes.data_as(c_int32_p) for k in ("nbr", "recv_begin", "recv_count", "send_ptr", "send_idx")],
	)

	out["send_idx"] = out["send_idx"][:info.n_send]
	out.update({name: getattr(info, name) for name, _ in PartitionInfo._fields_})
	return out


def csr_export(binding, c_matrix):
	"""(rowptr, col, val) of a CSR-kind bfm_matrix_t: structural pattern, original DOF order"""

	nnz = C.c_size_t()
	assert not binding.lib.bfmx_matrix_csr_export(C.byref(c_matrix), C.byref(nnz), None, None, None)

	rowptr = np.zeros(c_matrix.m + 1, np.uint64)
	col = np.zeros(nnz.value, np.uint64)
	val = np.zeros(nnz.value)

	assert not binding.lib.bfmx_matrix_csr_export(
		C.byref(c_matrix), C.byref(nnz), rowptr.ctypes.data_as(abi.c_size_t_p), col.ctypes.data_as(abi.c_size_t_p), val.ctypes.data_as(abi.c_double_p)
	)

	return rowptr, col, val


def csr_matrix(binding, rowptr, col, val):
	"""a CSR-kind bfm_matrix_t from scalar CSR arrays (bfmx_matrix_csr_create); keeps the arrays alive"""

	rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
	col = np.ascontiguousarray(col, dtype=np.uint64)
	val = np.ascontiguousarray(val, dtype=np.float64)
	matrix = abi.Matrix()

	assert not binding.lib.bfmx_matrix_csr_create(
		C.byref(matrix), C.byref(binding.state), len(rowptr) - 1,
		rowptr.ctypes.data_as(abi.c_size_t_p), col.ctypes.data_as(abi.c_size_t_p), val.ctypes.data_as(abi.c_double_p),
	)

	return matrix


def device_available(binding=None) -> bool:
	from .api import default_binding

	return bool((binding or default_binding()).lib.bfmx_device_available())


def last_stats(binding=None) -> dict:
	from .api import default_binding

	s = Stats()
	(binding or default_binding()).lib.bfmx_last_stats(C.byref(s))
	return s.as_dict()


def plate(nx: int, ny: int, lx: float = 4.0, ly: float = 1.0, kind: int = 3, with_edges: bool = False, binding=None):
	"""structured plate mesh generated by the library"""

	from .api import Mesh, default_binding

	binding = binding or default_binding()
	mesh = Mesh.__new__(Mesh)
	mesh.binding = binding
	mesh.c_mesh = abi.Mesh()
	assert not binding.lib.bfmx_mesh_plate(C.byref(mesh.c_mesh), C.byref(binding.state), nx, ny, lx, ly, kind, with_edges)
	mesh.dim = 2
	mesh.kind = kind
	return mesh


def internal_numbering(mesh):
	"""to_new (numpy int32) when bfm_sim_run would work on an internally renumbered copy of this mesh, else None"""

	to_new = np.zeros(mesh.c_mesh.n_nodes, np.int32)
	rv = mesh.binding.lib.bfmx_mesh_internal_numbering(C.byref(mesh.c_mesh), to_new.ctypes.data_as(c_int32_p))
	assert rv >= 0

	return to_new if rv == 1 else None


def pattern(mesh) -> dict:
	"""the symbolic plan of a mesh as numpy arrays (see bfmx_mesh_pattern_copy)"""

	lib = mesh.binding.lib
	sizes = [C.c_size_t() for _ in range(4)]
	assert not lib.bfmx_mesh_pattern_sizes(C.byref(mesh.c_mesh), *[C.byref(s) for s in sizes])
	n_slices, n_slots, n_blocks, n_ctr = [s.value for s in sizes]
	nb = mesh.c_mesh.n_nodes

	out = {
		"slice_off": np.zeros(n_slices + 1, np.int32),
		"row_len": np.zeros(nb, np.int32),
		"scol": np.zeros(n_slots, np.int32),
		"diag_pos": np.zeros(nb, np.int32),
		"ctr_ptr": np.zeros(n_slots + 1, np.int32),
		"ctr": np.zeros(max(n_ctr, 1), np.uint32),
	}

	assert not lib.bfmx_mesh_pattern_copy(
		C.byref(mesh.c_mesh),
		out["slice_off"].ctypes.data_as(c_int32_p), out["row_len"].ctypes.data_as(c_int32_p), out["scol"].ctypes.data_as(c_int32_p),
		out["diag_pos"].ctypes.data_as(c_int32_p), out["ctr_ptr"].ctypes.data_as(c_int32_p), out["ctr"].ctypes.data_as(c_uint32_p),
	)

	out["ctr"] = out["ctr"][:n_ctr]
	out.update(n_slices=n_slices, n_slots=n_slots, n_blocks=n_blocks, n_ctr=n_ctr, nb=nb)
	return out


def _sim_array(sims):
	arr = (_P(abi.Sim) * len(sims))()

	for i, sim in enumerate(sims):
		arr[i] = C.pointer(sim.c_sim)

	return arr


def sim_run_batch(sims):
	"""bfm_sim_run for many simulations at once: one assembly launch, one CTA per system (bfmx_sim_run_batch)"""

	lib = sims[0].binding.lib
	assert not lib.bfmx_sim_run_batch(_sim_array(sims), len(sims)), lib.bfmx_device_error()


class Job:
	"""staged pipeline of one instance (bfmx_job_*): create -> upload -> assemble -> solve -> download"""

	def __init__(self, sim, instance_index: int = 0):
		self.sim = sim
		self.lib = sim.binding.lib
		self.handle = C.c_void_p()
		assert not self.lib.bfmx_job_create(C.byref(self.handle), C.byref(sim.c_sim), instance_index), self.lib.bfmx_device_error()

	@classmethod
	def batch(cls, sims) -> "Job":
		"""one job for every instance of every simulation in `sims` (bfmx_job_create_batch)"""

		job = cls.__new__(cls)
		job.sim = list(sims)  # keeps them alive
		job.lib = sims[0].binding.lib
		job.handle = C.c_void_p()
		assert not job.lib.bfmx_job_create_batch(C.byref(job.handle), _sim_array(sims), len(sims)), job.lib.bfmx_device_error()
		return job

	def batch_status(self) -> list[dict]:
		out = []

		for i in range(self.lib.bfmx_job_batch_size(self.handle)):
			st = BatchStatus()
			assert not self.lib.bfmx_job_batch_status(self.handle, i, C.byref(st))
			out.append(st.as_dict())

		return out

	def __del__(self):
		try:
			if self.handle:
				self.lib.bfmx_job_destroy(self.handle)
		except Exception:
			pass

	def upload(self):
		assert not self.lib.bfmx_job_upload(self.handle)

	def assemble(self):
		assert not self.lib.bfmx_job_assemble(self.handle)

	def solve(self):
		assert not self.lib.bfmx_job_solve(self.handle)

	def download(self):
		assert not self.lib.bfmx_job_download(self.handle)

	def stats(self) -> dict:
		s = Stats()
		self.lib.bfmx_job_stats(self.handle, C.byref(s))
		return s.as_dict()

	def spmv_ms(self, reps: int = 20) -> float:
		ms = C.c_float()
		assert not self.lib.bfmx_job_spmv_time(self.handle, reps, C.byref(ms))
		return ms.value

	def read(self, rhs: bool = True, solution: bool = False):
		n = self.stats()["n_dofs"]
		b = np.zeros(n) if rhs else None
		x = np.zeros(n) if solution else None

		assert not self.lib.bfmx_job_read(
			self.handle,
			b.ctypes.data_as(abi.c_double_p) if rhs else None,
			x.ctypes.data_as(abi.c_double_p) if solution else None,
		)

		return b, x
