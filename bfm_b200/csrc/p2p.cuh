/*
 * p2p.cuh - exchanges over NVLink peer memory, done from inside the solver's own kernels.
 *
 * One process per GPU; at bfmx_dist_init every rank cudaMalloc's a "mailbox", exports it with CUDA IPC and
 * maps every peer's (dist.cu).  The small exchanges of a PCG iteration then need no collective launch:
 *
 *   scalars   the last CTA of a reducing kernel STORES its partial dot product into slot [me] of every
 *             peer's mailbox; the one-thread k_fold that follows spins until all slots of the round are
 *             there and folds them in rank order
 *   coarse    k_restrict stores W^T r (and the r.r share) into every peer; k_coarse_fold waits and folds
 *   mu        each rank applies ITS rows of E^-1 and stores its part of mu into every peer
 *   halo      k_halo_post packs the interface entries of a vector straight into the neighbour's staging area;
 *             k_halo_take waits for the neighbours' rounds and copies them into the ghost range (every distributed
 *             level of the multigrid hierarchy uses the same channel: the exchanges follow each other in stream order)
 *   gather    multigrid: every rank stores its slice of the first replicated level's right-hand side into every
 *             mailbox; k_mg_gather_take waits for all slices and copies the complete vector out
 *   panel     set-up only: the Gauss-Jordan inversion of the coarse operator is distributed by row blocks;
 *             the owner of a pivot block stores its row panel into every peer, and every rank acknowledges
 *             each finished step so that the two panel buffers can be reused safely
 *
 * Protocol (per channel): data stores, __threadfence_system(), then a store of the round number into
 * seq[buffer][me] on the receiver; the receiver polls its LOCAL seq words (volatile) and reads the data
 * with ld.cg (L2 is the point of coherence for peer writes).  Rounds are double-buffered by parity: a rank
 * can be at most one round ahead of a peer, because its next post needs that peer's current one.  Round
 * counters live on the device (Scalars) so that kernels skipped after convergence skip their rounds on
 * every rank alike; each solve starts from zeroed seq words behind a collective barrier.
 * Every spin has a time-out (10 s) that raises the mailbox's error word instead of hanging the GPU; once raised,
 * later waits of the same solve give up at once and the solve reports the failure.
 */
#pragma once

#include <cstdint>

constexpr int kP2pMaxRanks = BFMG_DIST_MAX_RANKS;
constexpr int kP2pScalarSlots = 8;

/* byte offsets inside a mailbox (identical on every rank) */
struct P2pLayout {
	size_t scalar_seq;  /* uint64 [2][kP2pMaxRanks] */
	size_t scalar_val;  /* double [2][kP2pMaxRanks][kP2pScalarSlots] */
	size_t coarse_seq;  /* uint64 [2][kP2pMaxRanks] */
	size_t mu_seq;      /* uint64 [2][kP2pMaxRanks] */
	size_t halo_seq;    /* uint64 [2][kP2pMaxRanks] */
	size_t gather_seq;  /* uint64 [2][kP2pMaxRanks]  multigrid: right-hand side of the first replicated level */
	size_t panel_seq;   /* uint64 [2][kP2pMaxRanks]  distributed Gauss-Jordan: pivot panels, by owner */
	size_t gj_done;     /* uint64 [kP2pMaxRanks]     ... steps every rank has finished (flow control) */
	size_t error;       /* int32 */
	size_t coarse_val;  /* double [2][world][coarse_cap] */
	size_t mu_val;      /* double [2][coarse_cap] */
	size_t halo_val;    /* double2 [2][world][halo_cap] */
	size_t panel_val;   /* double [2][32 * coarse_cap + 32 * 32 + 8]: row panel, inverse of the pivot block, bad flag */
	size_t gather_val;  /* double [2][gather_cap]: every rank writes its slice of the vector into every mailbox */
	size_t total;
	int32_t coarse_cap; /* doubles per rank in the coarse channel (n_c + 8 must fit) */
	int32_t halo_cap;   /* node entries (two doubles each) per neighbour in the halo channel */
	int32_t gather_cap; /* doubles in the gather channel */
};

struct P2p {
	char* box[kP2pMaxRanks]; /* box[me] is this rank's own mailbox, the others are IPC mappings */
	int32_t me;
	int32_t world;
	P2pLayout L;
};

#ifdef __CUDACC__

__device__ __forceinline__ void p2p_store_u64(uint64_t* p, uint64_t v) {
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t p2p_load_u64(uint64_t const* p) {
	uint64_t v;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void p2p_store_f64(double* p, double v) {
	asm volatile("st.volatile.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ uint64_t p2p_now_ns() {
	uint64_t t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

/* spin until the local word reaches `round`; false (and the error word set) after ~10 s, and at once when an
 * earlier wait of this solve has already failed - a broken solve drains quickly instead of timing out per kernel */
__device__ __forceinline__ bool p2p_wait(P2p const& X, uint64_t const* seq, uint64_t round) {
	if (p2p_load_u64(seq) >= round) {
		return true;
	}

	volatile int32_t* const error = (volatile int32_t*) (X.box[X.me] + X.L.error);

	if (*error) {
		return false;
	}

	uint64_t const t0 = p2p_now_ns();

	for (uint32_t spins = 0;; spins++) {
		if (p2p_load_u64(seq) >= round) {
			return true;
		}

		if ((spins & 1023) == 1023 && (*error || p2p_now_ns() - t0 > 10000000000ull)) {
			*error = 1;
			return false;
		}
	}
}

__device__ __forceinline__ uint64_t* p2p_seq(P2p const& X, int rank, size_t channel_off, uint64_t round, int slot) {
	return (uint64_t*) (X.box[rank] + channel_off) + (round & 1) * kP2pMaxRanks + slot;
}

/* after this thread's (and, through a ticket, its grid's) data stores: publish `round` of a channel to `rank` */
__device__ __forceinline__ void p2p_publish(P2p const& X, int rank, size_t channel_off, uint64_t round) {
	p2p_store_u64(p2p_seq(X, rank, channel_off, round, X.me), round);
}

#endif
