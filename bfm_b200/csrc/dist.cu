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
#include "p2p.cuh"

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

	/* NVLink peer-memory mailboxes (p2p.cuh); p2p_ok is false when IPC / peer access is unavailable
	 * or BFM_P2P=0, and the solver then uses the NCCL exchanges for everything */
	bool p2p_ok = false;
	void* mailbox = nullptr;
	P2p p2p = {};
	char p2p_why[256] = "";
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

size_t align_up(size_t v, size_t a) {
	return (v + a - 1) / a * a;
}

P2pLayout mailbox_layout(int world) {
	P2pLayout L = {};
	size_t at = 0;

	L.scalar_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.scalar_val = at, at += 2 * kP2pMaxRanks * kP2pScalarSlots * sizeof(double);
	L.coarse_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.mu_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.halo_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.gather_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.panel_seq = at, at += 2 * kP2pMaxRanks * sizeof(uint64_t);
	L.gj_done = at, at += kP2pMaxRanks * sizeof(uint64_t);
	L.error = at, at += 64;
	at = align_up(at, 256);

	L.coarse_cap = 3 * 6144 + 64;          /* n_c + 8 doubles per rank: up to 2048 x 16^(1/3) aggregates, and the binning may exceed its target */
	L.halo_cap = 1 << 16;                  /* interface nodes per neighbour */

	L.coarse_val = at, at += align_up((size_t) 2 * world * L.coarse_cap * sizeof(double), 256);
	L.mu_val = at, at += align_up((size_t) 2 * L.coarse_cap * sizeof(double), 256);
	L.halo_val = at, at += align_up((size_t) 2 * world * L.halo_cap * 2 * sizeof(double), 256);
	L.panel_val = at, at += align_up((size_t) 2 * (32 * (size_t) L.coarse_cap + 32 * 32 + 8) * sizeof(double), 256);
	L.gather_cap = 3 * 65536 + 64;         /* the first replicated level of the multigrid hierarchy: BFM_MG_REPLICATED_NODES nodes */
	L.gather_val = at, at += align_up((size_t) 2 * L.gather_cap * sizeof(double), 256);
	L.total = at;

	return L;
}

/* cudaMalloc a mailbox, exchange IPC handles over the communicator, map the peers' */
void setup_p2p() {
	char const* const env = getenv("BFM_P2P");

	D.p2p_ok = false;

	if (env != nullptr && atoi(env) == 0) {
		snprintf(D.p2p_why, sizeof D.p2p_why, "disabled by BFM_P2P=0");
		return;
	}

	P2pLayout const L = mailbox_layout(D.world);
	cudaIpcMemHandle_t mine;
	cudaIpcMemHandle_t all[kP2pMaxRanks];
	void* d_all = nullptr;

#define P2P_TRY(call)                                                                         \
	do {                                                                                      \
		cudaError_t const rc_ = (call);                                                       \
		if (rc_ != cudaSuccess) {                                                             \
			snprintf(D.p2p_why, sizeof D.p2p_why, "%s: %s", #call, cudaGetErrorString(rc_)); \
			cudaGetLastError();                                                               \
			failed = true;                                                                    \
		}                                                                                     \
	} while (0)

	bool failed = false;

	P2P_TRY(cudaMalloc(&D.mailbox, L.total));

	if (!failed) {
		P2P_TRY(cudaMemset(D.mailbox, 0, L.total));
		P2P_TRY(cudaIpcGetMemHandle(&mine, D.mailbox));
		P2P_TRY(cudaMalloc(&d_all, sizeof all));
	}

	/* every rank takes part in the all-gather whatever happened locally: a failed rank sends zeros */

	if (failed) {
		memset(&mine, 0, sizeof mine);
	}

	if (d_all == nullptr && cudaMalloc(&d_all, sizeof all) != cudaSuccess) {
		snprintf(D.p2p_why, sizeof D.p2p_why, "cudaMalloc for the handle exchange failed");
		return; /* cannot even take part; peers will time out in NCCL - as any allocation failure here would */
	}

	int ok_local = failed ? 0 : 1;

	cudaMemcpy((char*) d_all + (size_t) D.rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice);

	if (nccl_check(D.api.AllGather((char*) d_all + (size_t) D.rank * sizeof mine, d_all, sizeof mine, ncclChar, D.comm, bfmg_stream()), "handle exchange", __LINE__) < 0 || cudaStreamSynchronize(bfmg_stream()) != cudaSuccess) {
		snprintf(D.p2p_why, sizeof D.p2p_why, "handle exchange failed");
		cudaFree(d_all);
		return;
	}

	cudaMemcpy(all, d_all, sizeof mine * D.world, cudaMemcpyDeviceToHost);
	cudaFree(d_all);

	cudaIpcMemHandle_t zero;
	memset(&zero, 0, sizeof zero);

	for (int r = 0; r < D.world; r++) {
		if (memcmp(&all[r], &zero, sizeof zero) == 0) {
			ok_local = 0; /* some rank has no mailbox: nobody uses peer memory */
		}
	}

	D.p2p = P2p {};
	D.p2p.me = D.rank;
	D.p2p.world = D.world;
	D.p2p.L = L;

	for (int r = 0; r < D.world && ok_local; r++) {
		if (r == D.rank) {
			D.p2p.box[r] = (char*) D.mailbox;
			continue;
		}

		void* ptr = nullptr;
		cudaError_t const rc = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);

		if (rc != cudaSuccess) {
			snprintf(D.p2p_why, sizeof D.p2p_why, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(rc));
			cudaGetLastError();
			ok_local = 0;
			break;
		}

		D.p2p.box[r] = (char*) ptr;
	}

	/* all or nothing: agree over the communicator (reuse the first bytes of the mailbox-less path) */

	double* d_flag = nullptr;

	if (cudaMalloc((void**) &d_flag, sizeof(double) * (kP2pMaxRanks + 1)) != cudaSuccess) {
		return;
	}

	double const mine_ok = ok_local;
	double everyone[kP2pMaxRanks] = {};

	cudaMemcpy(d_flag, &mine_ok, sizeof mine_ok, cudaMemcpyHostToDevice);

	if (nccl_check(D.api.AllGather(d_flag, d_flag + 1, 1, ncclDouble, D.comm, bfmg_stream()), "p2p agreement", __LINE__) == 0 && cudaStreamSynchronize(bfmg_stream()) == cudaSuccess) {
		cudaMemcpy(everyone, d_flag + 1, sizeof(double) * D.world, cudaMemcpyDeviceToHost);

		bool all_ok = true;

		for (int r = 0; r < D.world; r++) {
			all_ok = all_ok && everyone[r] == 1;
		}

		D.p2p_ok = all_ok;

		if (!all_ok && D.p2p_why[0] == 0) {
			snprintf(D.p2p_why, sizeof D.p2p_why, "a peer could not map the mailboxes");
		}
	}

	cudaFree(d_flag);

#undef P2P_TRY
}

void teardown_p2p() {
	if (D.mailbox == nullptr) {
		return;
	}

	for (int r = 0; r < D.world; r++) {
		if (r != D.rank && D.p2p.box[r] != nullptr) {
			cudaIpcCloseMemHandle(D.p2p.box[r]);
		}
	}

	cudaFree(D.mailbox);
	cudaGetLastError();

	D.mailbox = nullptr;
	D.p2p = P2p {};
	D.p2p_ok = false;
}

/* sendbuf[i] = v[send_idx[i]] : the owned entries the neighbours ghost, grouped by neighbour */
__global__ void k_halo_pack(double2 const* __restrict__ v, int32_t const* __restrict__ send_idx, double2* __restrict__ sendbuf, int n) {
	pdl_sync();

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

	setup_p2p(); /* collective; on failure the NCCL exchanges stay in charge (bfmg_dist_p2p_status says why) */

	return 0;
}

int bfmg_dist_finalize(void) {
	if (D.comm != nullptr) {
		cudaStreamSynchronize(bfmg_stream());
		teardown_p2p();
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

char const* bfmg_dist_p2p_status(void) {
	return D.p2p_ok ? "" : (D.p2p_why[0] ? D.p2p_why : "not initialised");
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

/* ---- gpu_internal.cuh: peer-memory access for solver.cu --------------------------------------------- */

P2p const* bfmg_dist_p2p() {
	return D.p2p_ok ? &D.p2p : nullptr;
}

/* start of a solve: zero this rank's round words, then make sure every rank has done so before anyone
 * posts (one tiny all-gather as the barrier) */
int bfmg_dist_p2p_begin(int fits, int* all_fit) {
	*all_fit = 0;

	if (!D.p2p_ok) {
		return 0; /* decided collectively at init: the same on every rank */
	}

	/* the mu channel is idle until the first coarse round: its first doubles carry the votes */
	double* const votes = (double*) ((char*) D.mailbox + D.p2p.L.mu_val);
	double const mine = fits ? 1 : 0;
	double everyone[kP2pMaxRanks] = {};

	if (
		BFMG_CHECK(cudaMemsetAsync(D.mailbox, 0, D.p2p.L.coarse_val, bfmg_stream())) < 0 ||
		BFMG_CHECK(cudaMemcpyAsync(votes + D.rank, &mine, sizeof mine, cudaMemcpyHostToDevice, bfmg_stream())) < 0
	) {
		return -1;
	}

	D.collectives++;

	if (
		NCCL_CHECK(AllGather(votes + D.rank, votes, 1, ncclDouble, D.comm, bfmg_stream())) < 0 ||
		BFMG_CHECK(cudaMemcpyAsync(everyone, votes, sizeof(double) * D.world, cudaMemcpyDeviceToHost, bfmg_stream())) < 0 ||
		BFMG_CHECK(cudaStreamSynchronize(bfmg_stream())) < 0
	) {
		return -1;
	}

	int all = 1;

	for (int r = 0; r < D.world; r++) {
		all = all && everyone[r] == 1;
	}

	*all_fit = all;
	return 0;
}

/* end of a solve (stream synchronised): 1 when a peer wait timed out */
int bfmg_dist_p2p_failed() {
	if (!D.p2p_ok) {
		return 0;
	}

	int32_t err = 0;

	if (cudaMemcpy(&err, (char*) D.mailbox + D.p2p.L.error, sizeof err, cudaMemcpyDeviceToHost) != cudaSuccess) {
		cudaGetLastError();
		return 1;
	}

	return err != 0;
}
