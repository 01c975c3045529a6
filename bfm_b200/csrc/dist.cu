/*
 * dist.cu - the multi-GPU plumbing of libbfm (B200 build): one process per GPU, one NCCL communicator
 * over NVLink 5 / NVSwitch, everything enqueued on the library's own stream.
 *
 * The reference has no distributed code at all; this is the row-partitioned
 * CG the north star asks for.  Three exchanges exist on the solve path:
 *
 *   halo      before every SpMV, the entries of the gathered vector at the nodes a neighbour ghosts:
 *             one pack kernel + one grouped ncclSend/ncclRecv per neighbour, received straight into
 *             the ghost range of the vector (ghosts of one owner are contiguous, partition.c)
 *   scalar    the two dot products of a CG iteration: ncclAllGather of one double per rank, folded in
 *             rank order by every rank (solver.cu) - all ranks get the SAME bits, so the convergence
 *             decision and with it the sequence of collectives cannot diverge between ranks
 *   gather    once per solve, the owned displacement blocks into the full vector on every rank
 *
 * NCCL is loaded with dlopen at bfmx_dist_init time, so single-GPU users of libbfm.so do not need it
 * (and a process that already carries torch's NCCL shares that copy).
 */
#include "gpu_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

namespace {

struct Nccl {
	void* handle = nullptr;

	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	char const* (*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*Send)(void const*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllGather)(void const*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(void const*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
};

struct Dist {
	Nccl api;
	ncclComm_t comm = nullptr;
	int rank = 0;
	int world = 1;
	size_t collectives = 0;
};

Dist D;

bool load_nccl() {
	if (D.api.handle != nullptr) {
		return true;
	}

	char const* const names[] = {getenv("BFM_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};

	for (char const* name : names) {
		if (name != nullptr && (D.api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL)) != nullptr) {
			break;
		}
	}

	if (D.api.handle == nullptr) {
		bfmg_set_error("cannot load NCCL (libnccl.so.2): %s", dlerror());
		return false;
	}

#define RESOLVE(field, symbol)                                                        \
	if ((*(void**) &D.api.field = dlsym(D.api.handle, symbol)) == nullptr) {          \
		bfmg_set_error("NCCL library lacks %s", symbol);                              \
		dlclose(D.api.handle);                                                        \
		D.api.handle = nullptr;                                                       \
		return false;                                                                 \
	}

	RESOLVE(GetUniqueId, "ncclGetUniqueId")
	RESOLVE(CommInitRank, "ncclCommInitRank")
	RESOLVE(CommDestroy, "ncclCommDestroy")
	RESOLVE(GetErrorString, "ncclGetErrorString")
	RESOLVE(GroupStart, "ncclGroupStart")
	RESOLVE(GroupEnd, "ncclGroupEnd")
	RESOLVE(Send, "ncclSend")
	RESOLVE(Recv, "ncclRecv")
	RESOLVE(AllGather, "ncclAllGather")
	RESOLVE(Broadcast, "ncclBroadcast")
	RESOLVE(GetVersion, "ncclGetVersion")

#undef RESOLVE

	return true;
}

int nccl_check(ncclResult_t rc, char const* what, int line) {
	if (rc == ncclSuccess) {
		return 0;
	}

	bfmg_set_error("%s: %s (dist.cu:%d)", what, D.api.GetErrorString(rc), line);
	return -1;
}

#define NCCL_CHECK(call) nccl_check(D.api.call, #call, __LINE__)

/* sendbuf[i] = v[send_idx[i]] : the owned entries the neighbours ghost, grouped by neighbour */
__global__ void k_halo_pack(double2 const* __restrict__ v, int32_t const* __restrict__ send_idx, double2* __restrict__ sendbuf, int n) {
	int const i = blockIdx.x * blockDim.x + threadIdx.x;

	if (i < n) {
		sendbuf[i] = v[send_idx[i]];
	}
}

} // namespace

extern "C" {

int bfmg_dist_unique_id(void* id) {
	if (!bfmg_ready() || !load_nccl()) {
		return -1;
	}

	static_assert(sizeof(ncclUniqueId) == BFMG_DIST_ID_BYTES, "bfmx_dist_unique_id hands out NCCL_UNIQUE_ID_BYTES bytes");

	ncclUniqueId uid;

	if (NCCL_CHECK(GetUniqueId(&uid)) < 0) {
		return -1;
	}

	memcpy(id, &uid, sizeof uid);
	return 0;
}

int bfmg_dist_init(int rank, int world, void const* id) {
	if (!bfmg_ready()) {
		return -1;
	}

	if (world < 1 || rank < 0 || rank >= world || world > BFMG_DIST_MAX_RANKS) {
		bfmg_set_error("bad rank %d / world %d (at most %d ranks)", rank, world, BFMG_DIST_MAX_RANKS);
		return -1;
	}

	if (D.comm != nullptr) {
		bfmg_set_error("bfmx_dist_init called twice");
		return -1;
	}

	if (world == 1) {
		D.rank = 0;
		D.world = 1;
		return 0;
	}

	if (!load_nccl()) {
		return -1;
	}

	ncclUniqueId uid;
	memcpy(&uid, id, sizeof uid);

	if (NCCL_CHECK(CommInitRank(&D.comm, world, uid, rank)) < 0) {
		D.comm = nullptr;
		return -1;
	}

	D.rank = rank;
	D.world = world;

	return 0;
}

int bfmg_dist_finalize(void) {
	if (D.comm != nullptr) {
		cudaStreamSynchronize(bfmg_stream());
		D.api.CommDestroy(D.comm);
		D.comm = nullptr;
	}

	D.rank = 0;
	D.world = 1;

	return 0;
}

int bfmg_dist_world(void) {
	return D.world;
}

int bfmg_dist_rank(void) {
	return D.rank;
}

size_t bfmg_dist_collectives(void) {
	return D.collectives;
}

int bfmg_dist_halo(bfmg_halo_t const* halo, double* d_vec, double* d_sendbuf) {
	if (D.comm == nullptr || halo->n_nbr == 0) {
		return 0;
	}

	double2* const v = (double2*) d_vec;
	double2* const sendbuf = (double2*) d_sendbuf;

	if (halo->n_send > 0 && BFMG_LAUNCH(k_halo_pack, (halo->n_send + kBlock - 1) / kBlock, kBlock, 0, v, halo->d_send_idx, sendbuf, halo->n_send) < 0) {
		return -1;
	}

	if (NCCL_CHECK(GroupStart()) < 0) {
		return -1;
	}

	int rv = 0;

	for (int i = 0; i < halo->n_nbr; i++) {
		int const n_send = halo->send_ptr[i + 1] - halo->send_ptr[i];

		if (n_send > 0) {
			rv |= NCCL_CHECK(Send(sendbuf + halo->send_ptr[i], 2 * (size_t) n_send, ncclDouble, halo->nbr[i], D.comm, bfmg_stream()));
		}

		if (halo->recv_count[i] > 0) {
			rv |= NCCL_CHECK(Recv(v + halo->recv_begin[i], 2 * (size_t) halo->recv_count[i], ncclDouble, halo->nbr[i], D.comm, bfmg_stream()));
		}
	}

	rv |= NCCL_CHECK(GroupEnd());
	D.collectives++;

	return rv;
}

int bfmg_dist_allgather_f64(double const* d_send, double* d_recv, int count) {
	if (D.comm == nullptr) {
		return BFMG_CHECK(cudaMemcpyAsync(d_recv, d_send, (size_t) count * sizeof(double), cudaMemcpyDeviceToDevice, bfmg_stream()));
	}

	D.collectives++;
	return NCCL_CHECK(AllGather(d_send, d_recv, (size_t) count, ncclDouble, D.comm, bfmg_stream()));
}

int bfmg_dist_gather_blocks(double const* d_owned, size_t const* first_node, double* d_global) {
	if (D.comm == nullptr) {
		return BFMG_CHECK(cudaMemcpyAsync(d_global, d_owned, 2 * (first_node[1] - first_node[0]) * sizeof(double), cudaMemcpyDeviceToDevice, bfmg_stream()));
	}

	if (NCCL_CHECK(GroupStart()) < 0) {
		return -1;
	}

	int rv = 0;

	for (int r = 0; r < D.world; r++) {
		size_t const count = 2 * (first_node[r + 1] - first_node[r]);

		if (count > 0) {
			rv |= NCCL_CHECK(Broadcast(d_owned, d_global + 2 * first_node[r], count, ncclDouble, r, D.comm, bfmg_stream()));
		}
	}

	rv |= NCCL_CHECK(GroupEnd());
	D.collectives++;

	return rv;
}

} // extern "C"
