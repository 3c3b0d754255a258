/*
 * dist.cu -- clo_dist_*: the three operators sharded over the GPUs of one box, behind the
 * C ABI (include/cl_ops/clo_b200.h).  The reference is single-device
 * (/root/reference/src/cl_ops/sort/clo_sort_abstract.c:335), so there is no counterpart; the
 * partitioning follows SURVEY.md section 8(e).
 *
 * Data never goes through the caller's communicator: the sample sort's exchange is
 * clo_partition_scatter writing into peer memory (partition.cu).  The communicator carries the
 * sample rows, the bucket sizes and one barrier per sort.
 *
 * Splitters are chosen by this file's own kernel: every gathered sample computes its rank in
 * the (key, gathered position) order by counting -- the O(S^2) selection of the reference's
 * gselect (clo_sort_gselect.cl:27-55) applied to S <= 64 * world^2 samples -- and the samples
 * whose ranks are the regular picks k * S / world become the splitters.  No sort, no host.
 */
#include "clo_internal.h"
#include "device_utils.cuh"

#include <cl_ops/clo_b200.h>

#include <cstring>
#include <vector>

using namespace clo;

namespace {

const int DIST_MAX_WORLD = 16;
const int DIST_PHASES = 7;

__device__ __forceinline__ u64 dist_load_key(const void* keys, size_t i, int kb) {
	return kb == 4 ? (u64) reinterpret_cast<const u32*>(keys)[i] : reinterpret_cast<const u64*>(keys)[i];
}

/* row = [numel, s, key of sample 0 .. cap-1]; sample j sits at position (j * numel) / s */
__global__ void clo_dist_sample(const void* __restrict__ keys, u64 numel, u32 cap, int kb, u64* __restrict__ row) {
	const u32 s = numel < cap ? (u32) numel : cap;
	const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j == 0) { row[0] = numel; row[1] = s; }
	if (j < cap) row[2 + j] = j < s ? dist_load_key(keys, (size_t) (((u64) j * numel) / s), kb) : 0ull;
}

/* all: [world][2 + cap] gathered rows.  One thread per sample slot.  Writes the world-1 splitters
 * as (key, index RELATIVE to this rank's first element, clamped at 0): with them the partition
 * kernels run with gidx0 = 0 and the host never needs the global offsets.  info[0] = this rank's
 * global offset, info[1] = total number of elements. */
__global__ void __launch_bounds__(256)
clo_dist_splitters(const u64* __restrict__ all, u32 row_w, u32 world, u32 cap, u32 rank, u64 gidx0_given,
		int kb, void* __restrict__ spl_keys, u64* __restrict__ spl_idx, u64* __restrict__ info) {
	/* a CTA ranks 32 samples; its 8 warps each compare them with one eighth of every 2048-slot tile */
	constexpr int TILE = 2048;
	__shared__ u64 s_keys[TILE];
	__shared__ unsigned char s_valid[TILE];
	__shared__ u32 s_part[8][32];
	__shared__ u64 s_g0[DIST_MAX_WORLD + 1];
	__shared__ u32 s_cnt[DIST_MAX_WORLD + 1];
	const u32 lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		u64 g = 0; u32 c = 0;
		for (u32 r = 0; r < world; ++r) { s_g0[r] = g; s_cnt[r] = c; g += all[(size_t) r * row_w]; c += (u32) all[(size_t) r * row_w + 1]; }
		s_g0[world] = g; s_cnt[world] = c;
	}
	__syncthreads();
	const u32 total = s_cnt[world];
	const u32 slots = world * cap;
	const u32 me = blockIdx.x * 32 + lane;
	const u32 my_r = me / cap, my_j = me % cap;
	const bool valid = me < slots && my_j < (u32) all[(size_t) my_r * row_w + 1];
	const u64 my_key = valid ? all[(size_t) my_r * row_w + 2 + my_j] : 0ull;
	u32 below = 0;
	for (u32 base = 0; base < slots; base += TILE) {
		__syncthreads();
		for (u32 i = threadIdx.x; i < (u32) TILE; i += 256) {
			const u32 o = base + i;
			const u32 r = o / cap, j = o % cap;
			const bool v = o < slots && j < (u32) all[(size_t) r * row_w + 1];
			s_keys[i] = v ? all[(size_t) r * row_w + 2 + j] : ~0ull;
			s_valid[i] = v ? 1 : 0;                    /* an empty slot never counts */
		}
		__syncthreads();
		if (valid) {
			const u32 t0 = slice * (TILE / 8);
#pragma unroll 8
			for (u32 t = t0; t < t0 + TILE / 8; ++t) {
				const u64 k = s_keys[t];
				below += (s_valid[t] && (k < my_key || (k == my_key && base + t < me))) ? 1u : 0u;
			}
		}
	}
	s_part[slice][lane] = below;
	__syncthreads();
	if (blockIdx.x == 0 && threadIdx.x == 0) { info[0] = gidx0_given != ~0ull ? gidx0_given : s_g0[rank]; info[1] = s_g0[world]; }
	if (slice != 0 || !valid || total == 0) return;
	below = 0;
#pragma unroll
	for (int w = 0; w < 8; ++w) below += s_part[w][lane];
	const u64 my_g0 = gidx0_given != ~0ull ? gidx0_given : s_g0[rank];
	for (u32 k = 1; k < world; ++k) {
		const u32 pick = (u32) (((u64) k * total) / world);
		if (below == pick) {
			const u64 n_r = all[(size_t) my_r * row_w], s_r = all[(size_t) my_r * row_w + 1];
			const u64 gidx = s_g0[my_r] + ((u64) my_j * n_r) / s_r;
			if (kb == 4) reinterpret_cast<u32*>(spl_keys)[k - 1] = (u32) my_key;
			else reinterpret_cast<u64*>(spl_keys)[k - 1] = my_key;
			spl_idx[k - 1] = gidx > my_g0 ? gidx - my_g0 : 0ull;
		}
	}
}

/* M[src][dst] bucket sizes -> first slot of my bucket in every destination, fits flag, and for
 * the host (mapped memory): what I receive from every source, what I send to every destination,
 * the flag and the largest slice any rank receives */
__global__ void clo_dist_slots(const u64* __restrict__ M, u32 mstride, u32 world, u32 rank, u64 capacity,
		u64* __restrict__ first_slot, int* __restrict__ ok, volatile u64* __restrict__ host) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	int fits = 1;
	u64 worst = 0;
	for (u32 q = 0; q < world; ++q) {
		u64 before = 0, col = 0;
		for (u32 s = 0; s < world; ++s) { const u64 m = M[s * mstride + q]; if (s < rank) before += m; col += m; }
		first_slot[q] = before;
		if (col > capacity) fits = 0;
		if (col > worst) worst = col;
	}
	*ok = fits;
	for (u32 s = 0; s < world; ++s) { host[s] = M[s * mstride + rank]; host[DIST_MAX_WORLD + s] = M[rank * mstride + s]; }
	host[2 * DIST_MAX_WORLD] = (u64) fits;
	host[2 * DIST_MAX_WORLD + 1] = worst;
}

/* ---- control plane over peer memory.
 * The bookkeeping collectives (sample rows, bucket sizes, per-GPU totals, "my writes have landed")
 * are a few hundred bytes each; as NCCL calls they cost 40-150 us apiece at 8 GPUs, most of it
 * launch and host latency.  Every rank owns a small CONTROL buffer that all peers map (CUDA IPC):
 *   bcast: one CTA copies this rank's words into slot [rank] of every peer's buffer, fences at
 *          system scope, and stores the call's epoch into flag [rank] of every peer;
 *   wait:  one CTA spins (bounded) until its own flags of all ranks carry the epoch.
 * A kernel boundary separates the wait from the consumer, so the consumer's loads cannot hit
 * stale L1 lines.  Slots are reused by the next call only after that call's own flags, which a
 * rank publishes after it consumed the previous ones. */
const int CTL_ROW_MAX = 2 + 64 * DIST_MAX_WORLD;          /* words of a sample row */
const size_t CTL_ROWS = 0;                                  /* [16][CTL_ROW_MAX] */
const size_t CTL_SIZES = CTL_ROWS + (size_t) DIST_MAX_WORLD * CTL_ROW_MAX;    /* [16][16] */
const size_t CTL_TOTALS = CTL_SIZES + DIST_MAX_WORLD * DIST_MAX_WORLD;        /* [2][16]: by epoch parity (a scan has no closing barrier) */
const size_t CTL_FLAGS = CTL_TOTALS + 2 * DIST_MAX_WORLD;                     /* [4][16] */
const size_t CTL_WORDS = CTL_FLAGS + 4 * DIST_MAX_WORLD;
enum { CTL_F_ROWS = 0, CTL_F_SIZES = 1, CTL_F_DONE = 2, CTL_F_TOTALS = 3 };

__global__ void __launch_bounds__(1024)
clo_dist_bcast(const u64* __restrict__ src, u32 nwords, u64* const* __restrict__ ctl_ptrs, u32 world, u32 rank,
		size_t slot_off, u32 flag, u64 epoch) {
	for (u32 p = 0; p < world; ++p) {
		u64* dst = ctl_ptrs[p] + slot_off;
		for (u32 i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x < world) {
		volatile u64* f = ctl_ptrs[threadIdx.x] + CTL_FLAGS + (size_t) flag * DIST_MAX_WORLD + rank;
		*f = epoch;
	}
}

__global__ void clo_dist_wait(const u64* __restrict__ ctl, u32 world, u32 flag, u64 epoch, int* __restrict__ err) {
	if (threadIdx.x < world) {
		const volatile u64* f = ctl + CTL_FLAGS + (size_t) flag * DIST_MAX_WORLD + threadIdx.x;
		const long long t0 = clock64();
		while (*f < epoch) {
			if (clock64() - t0 > 60000000000ll) { atomicExch(err, 1); break; }     /* ~30 s: a peer never arrived (ranks may be seconds apart on a first call) */
			__nanosleep(200);
		}
	}
	__threadfence_system();
}

/* carry-in of rank `rank` from the gathered per-rank totals, in the scan's sum arithmetic */
template <typename SumT, typename AccT>
__global__ void clo_dist_carry(const SumT* __restrict__ totals, u32 rank, SumT* __restrict__ carry) {
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	AccT acc = AccT(0);
	for (u32 r = 0; r < rank; ++r) acc += static_cast<AccT>(totals[r * (8 / sizeof(SumT))]);
	*carry = static_cast<SumT>(acc);
}

} // namespace

struct clo_dist {
	CCLContext* ctx;
	CloDistComm comm;
	/* sort */
	CloSort* sorter = nullptr;
	CloType key_type = CLO_UINT;
	int kb = 4;
	bool with_payload = false;
	size_t capacity = 0;
	u32 cap = 0;                                  /* samples per rank */
	CCLBuffer* recv_k = nullptr;
	CCLBuffer* recv_p = nullptr;
	std::vector<CCLBuffer*> peers;                /* imported buffers (to release) */
	/* device work area (one allocation) and buffer views on it */
	void* work = nullptr;
	CCLBuffer *b_splk = nullptr, *b_spli = nullptr, *b_counts = nullptr, *b_first = nullptr, *b_ok = nullptr,
		*b_ptrs_k = nullptr, *b_ptrs_p = nullptr;
	u64 *d_row = nullptr, *d_all = nullptr, *d_M = nullptr, *d_info = nullptr, *d_tot = nullptr, *d_tots = nullptr, *d_carry = nullptr;
	u64* h_vec = nullptr;                         /* pinned, mapped: [recv x16 | send x16 | fits | worst] */
	u64* h_vec_dev = nullptr;
	/* control plane over peer memory (NULL: the communicator's device collectives are used) */
	CCLBuffer* ctl = nullptr;
	std::vector<CCLBuffer*> ctl_peers;
	u64** d_ctl_ptrs = nullptr;                   /* device array [16] of every rank's control buffer */
	int* d_err = nullptr;
	u64 epoch = 0;
	bool timing = false;
	cudaEvent_t ev[DIST_PHASES + 1] = {};
	bool ev_valid = false;
	u64 last_sent[DIST_MAX_WORLD] = {}, last_recv[DIST_MAX_WORLD] = {};
};

static void dist_host_barrier(CloDist* d) {
	if (d->comm.world > 1 && d->comm.all_gather_host) {
		unsigned char a = 0, all[DIST_MAX_WORLD];
		d->comm.all_gather_host(d->comm.user, &a, all, 1);
	}
}

/* collective when receive buffers exist: nobody may still be writing into a buffer that is about
 * to be unmapped, and nobody may free a buffer a peer still has mapped */
static void dist_release(CloDist* d) {
	CloDeviceGuard g(d->ctx->dev.ordinal);
	const bool shared = d->recv_k != nullptr;
	if (shared) { cudaDeviceSynchronize(); dist_host_barrier(d); }
	for (CCLBuffer* b : d->peers) ccl_buffer_destroy(b);
	d->peers.clear();
	if (shared) dist_host_barrier(d);
	CCLBuffer** views[] = { &d->b_splk, &d->b_spli, &d->b_counts, &d->b_first, &d->b_ok, &d->b_ptrs_k, &d->b_ptrs_p, &d->recv_k, &d->recv_p };
	for (CCLBuffer** v : views) { if (*v) ccl_buffer_destroy(*v); *v = nullptr; }
	if (d->work) cudaFree(d->work);
	d->work = nullptr;
	if (d->h_vec) cudaFreeHost(d->h_vec);
	d->h_vec = nullptr;
	if (d->sorter) clo_sort_destroy(d->sorter);
	d->sorter = nullptr;
	for (cudaEvent_t& e : d->ev) { if (e) cudaEventDestroy(e); e = nullptr; }
}

extern "C" CloDist* clo_dist_new(CCLContext* ctx, const CloDistComm* comm, GError** err) {
	if (!ctx || !comm || comm->world < 1 || comm->world > (cl_uint) DIST_MAX_WORLD || comm->rank >= comm->world) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_new: a communicator with 1..%d ranks is required", DIST_MAX_WORLD);
		return NULL;
	}
	const char* cm = getenv("CLO_DIST_CTRL");
	const bool want_peer = comm->world > 1 && comm->all_gather_host && !(cm && strcmp(cm, "nccl") == 0);
	if (comm->world > 1 && !want_peer && (!comm->all_gather_dev || !comm->barrier_dev)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_new: the communicator needs all_gather_host (control plane over peer memory) or all_gather_dev + barrier_dev");
		return NULL;
	}
	CloDist* d = new clo_dist();
	d->ctx = ctx;
	d->comm = *comm;
	ccl_context_ref(ctx);
	CloDeviceGuard g(ctx->dev.ordinal);
	/* scan work words (the sort's area comes with clo_dist_sort_setup) */
	void* p = nullptr;
	if (clo_cuda_failed(cudaMalloc(&p, 8 * (4 + DIST_MAX_WORLD)), err, "clo_dist_new")) { ccl_context_unref(ctx); delete d; return NULL; }
	cudaMemset(p, 0, 8 * (4 + DIST_MAX_WORLD));
	d->d_tot = (u64*) p; d->d_carry = d->d_tot + 1; d->d_tots = d->d_tot + 2;
	d->d_err = (int*) (d->d_tot + 2 + DIST_MAX_WORLD);
	clo_handle_add(d);
	if (want_peer) {
		/* collective: every rank allocates its control buffer, the handles go round, every peer's is mapped */
		const u32 P = comm->world, r = comm->rank;
		bool ok = true;
		d->ctl = ccl_buffer_new(ctx, 0, CTL_WORDS * 8, NULL, err);
		ok = d->ctl != NULL;
		if (ok) ok = !clo_cuda_failed(cudaMemset(d->ctl->ptr, 0, CTL_WORDS * 8), err, "clo_dist control buffer");
		unsigned char mine[64] = {}, all[DIST_MAX_WORLD * 64];
		if (ok) ok = clo_b200_ipc_export(d->ctl, mine, err) != 0;
		/* the exchange is entered even after a local failure, so that the peers do not hang */
		unsigned char flag_mine = ok ? 1 : 0, flags[DIST_MAX_WORLD];
		if (comm->all_gather_host(comm->user, mine, all, 64) != 0 || comm->all_gather_host(comm->user, &flag_mine, flags, 1) != 0) ok = false;
		for (u32 i = 0; ok && i < P; ++i) if (!flags[i]) ok = false;
		u64 h[DIST_MAX_WORLD] = {};
		for (u32 i = 0; ok && i < P; ++i) {
			if (i == r) { h[i] = (u64) (uintptr_t) d->ctl->ptr; continue; }
			CCLBuffer* b = clo_b200_ipc_import(ctx, all + (size_t) i * 64, CTL_WORDS * 8, err);
			if (!b) { ok = false; break; }
			d->ctl_peers.push_back(b);
			h[i] = (u64) (uintptr_t) b->ptr;
		}
		if (ok) ok = !clo_cuda_failed(cudaMalloc((void**) &d->d_ctl_ptrs, sizeof(h)), err, "clo_dist control pointers");
		if (ok) ok = !clo_cuda_failed(cudaMemcpy(d->d_ctl_ptrs, h, sizeof(h), cudaMemcpyHostToDevice), err, "clo_dist control pointers");
		unsigned char a = ok ? 1 : 0, b2[DIST_MAX_WORLD];
		comm->all_gather_host(comm->user, &a, b2, 1);          /* everyone is mapped (or somebody failed) */
		for (u32 i = 0; i < P; ++i) if (!b2[i]) ok = false;
		if (!ok) {
			if (err && !*err) g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_new: setting up the peer control buffers failed on some rank");
			/* local clean-up only: the ranks may disagree about what exists */
			for (CCLBuffer* b : d->ctl_peers) ccl_buffer_destroy(b);
			if (d->ctl) ccl_buffer_destroy(d->ctl);
			if (d->d_ctl_ptrs) cudaFree(d->d_ctl_ptrs);
			cudaFree(d->d_tot);
			clo_handle_remove(d);
			ccl_context_unref(ctx);
			delete d;
			return NULL;
		}
	}
	return d;
}

extern "C" void clo_dist_destroy(CloDist* d) {
	if (!d || !clo_handle_remove(d)) return;
	dist_release(d);
	{
		CloDeviceGuard g(d->ctx->dev.ordinal);
		if (d->ctl) {
			/* nobody may still be writing flags into a control buffer that is about to go */
			cudaDeviceSynchronize();
			dist_host_barrier(d);
			for (CCLBuffer* b : d->ctl_peers) ccl_buffer_destroy(b);
			d->ctl_peers.clear();
			dist_host_barrier(d);
			ccl_buffer_destroy(d->ctl);
			d->ctl = nullptr;
			if (d->d_ctl_ptrs) cudaFree(d->d_ctl_ptrs);
		}
		cudaFree(d->d_tot);
	}
	ccl_context_unref(d->ctx);
	delete d;
}

extern "C" cl_bool clo_dist_sort_setup(CloDist* d, CloType key_type, size_t capacity, cl_bool with_payload, GError** err) {
	if (!d || !clo_handle_alive(d)) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort_setup: invalid handle"); return CL_FALSE; }
	const size_t kb = clo_type_sizeof(key_type);
	if ((kb != 4 && kb != 8) || key_type == CLO_FLOAT || key_type == CLO_DOUBLE) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort_setup: integer keys of 4 or 8 bytes");
		return CL_FALSE;
	}
	if (d->comm.world > 1 && !d->comm.all_gather_host) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort_setup: the communicator has no all_gather_host");
		return CL_FALSE;
	}
	if (d->recv_k) dist_release(d);
	const u32 P = d->comm.world, r = d->comm.rank;
	CloDeviceGuard g(d->ctx->dev.ordinal);
	d->key_type = key_type; d->kb = (int) kb; d->with_payload = with_payload != 0; d->capacity = capacity;
	d->cap = 64 * P;
	/* satradix sorts raw key bits ascending whatever the signedness: sort as the unsigned type */
	CloType st_type = kb == 4 ? CLO_UINT : CLO_ULONG;
	d->sorter = clo_sort_new("satradix", NULL, d->ctx, &st_type, NULL, NULL, NULL, NULL, err);
	if (!d->sorter) return CL_FALSE;
	d->recv_k = ccl_buffer_new(d->ctx, 0, (capacity ? capacity : 1) * kb, NULL, err);
	if (!d->recv_k) return CL_FALSE;
	if (with_payload) {
		d->recv_p = ccl_buffer_new(d->ctx, 0, (capacity ? capacity : 1) * 4, NULL, err);
		if (!d->recv_p) return CL_FALSE;
	}
	/* work area */
	const size_t row_w = 2 + d->cap;
	const size_t words = row_w + (size_t) P * row_w + 5 * DIST_MAX_WORLD + (size_t) P * P + 8;
	if (clo_cuda_failed(cudaMalloc(&d->work, words * 8), err, "clo_dist work area")) return CL_FALSE;
	if (clo_cuda_failed(cudaMemset(d->work, 0, words * 8), err, "clo_dist work area")) return CL_FALSE;
	u64* w = (u64*) d->work;
	d->d_row = w; w += row_w;
	d->d_all = w; w += (size_t) P * row_w;
	u64* splk = w; w += DIST_MAX_WORLD;
	u64* spli = w; w += DIST_MAX_WORLD;
	u64* counts = w; w += DIST_MAX_WORLD;
	u64* first = w; w += DIST_MAX_WORLD;
	u64* ptrs = w; w += DIST_MAX_WORLD;             /* keys [0..16) ; payload pointers share the tail below */
	d->d_M = w; w += (size_t) P * P;
	d->d_info = w; w += 2;
	u64* okw = w; w += 1;
	(void) w;
	void* ptrs_p = nullptr;
	if (with_payload) {
		if (clo_cuda_failed(cudaMalloc(&ptrs_p, DIST_MAX_WORLD * 8), err, "clo_dist pointer table")) return CL_FALSE;
	}
	d->b_splk = ccl_buffer_new_wrap(d->ctx, splk, DIST_MAX_WORLD * 8, err);
	d->b_spli = ccl_buffer_new_wrap(d->ctx, spli, DIST_MAX_WORLD * 8, err);
	d->b_counts = ccl_buffer_new_wrap(d->ctx, counts, DIST_MAX_WORLD * 8, err);
	d->b_first = ccl_buffer_new_wrap(d->ctx, first, DIST_MAX_WORLD * 8, err);
	d->b_ok = ccl_buffer_new_wrap(d->ctx, okw, 8, err);
	d->b_ptrs_k = ccl_buffer_new_wrap(d->ctx, ptrs, DIST_MAX_WORLD * 8, err);
	if (with_payload) {
		d->b_ptrs_p = ccl_buffer_new_wrap(d->ctx, ptrs_p, DIST_MAX_WORLD * 8, err);
		if (d->b_ptrs_p) d->b_ptrs_p->owns = true;     /* freed with the view */
	}
	if (!d->b_splk || !d->b_spli || !d->b_counts || !d->b_first || !d->b_ok || !d->b_ptrs_k || (with_payload && !d->b_ptrs_p)) return CL_FALSE;
	if (clo_cuda_failed(cudaHostAlloc((void**) &d->h_vec, (2 * DIST_MAX_WORLD + 2) * 8, cudaHostAllocMapped), err, "clo_dist host words")) return CL_FALSE;
	if (clo_cuda_failed(cudaHostGetDevicePointer((void**) &d->h_vec_dev, d->h_vec, 0), err, "clo_dist host words")) return CL_FALSE;
	/* exchange the IPC handles of the receive buffers; map every peer's */
	unsigned char mine[128] = {}, all[DIST_MAX_WORLD * 128];
	if (P > 1) {
		if (!clo_b200_ipc_export(d->recv_k, mine, err)) return CL_FALSE;
		if (with_payload && !clo_b200_ipc_export(d->recv_p, mine + 64, err)) return CL_FALSE;
		if (d->comm.all_gather_host(d->comm.user, mine, all, 128) != 0) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_sort_setup: all_gather_host failed");
			return CL_FALSE;
		}
	}
	u64 hk[DIST_MAX_WORLD] = {}, hp[DIST_MAX_WORLD] = {};
	for (u32 i = 0; i < P; ++i) {
		if (i == r) { hk[i] = (u64) (uintptr_t) d->recv_k->ptr; hp[i] = with_payload ? (u64) (uintptr_t) d->recv_p->ptr : 0; continue; }
		CCLBuffer* bk = clo_b200_ipc_import(d->ctx, all + (size_t) i * 128, d->recv_k->size, err);
		if (!bk) return CL_FALSE;
		d->peers.push_back(bk);
		hk[i] = (u64) (uintptr_t) bk->ptr;
		if (with_payload) {
			CCLBuffer* bp = clo_b200_ipc_import(d->ctx, all + (size_t) i * 128 + 64, d->recv_p->size, err);
			if (!bp) return CL_FALSE;
			d->peers.push_back(bp);
			hp[i] = (u64) (uintptr_t) bp->ptr;
		}
	}
	if (clo_cuda_failed(cudaMemcpy(ptrs, hk, sizeof(hk), cudaMemcpyHostToDevice), err, "clo_dist pointer table")) return CL_FALSE;
	if (with_payload && clo_cuda_failed(cudaMemcpy(ptrs_p, hp, sizeof(hp), cudaMemcpyHostToDevice), err, "clo_dist pointer table")) return CL_FALSE;
	for (cudaEvent_t& e : d->ev) if (!e && clo_cuda_failed(cudaEventCreate(&e), err, "clo_dist events")) return CL_FALSE;
	if (P > 1) { unsigned char a = 0, b[DIST_MAX_WORLD]; d->comm.all_gather_host(d->comm.user, &a, b, 1); }   /* everyone is mapped */
	return CL_TRUE;
}

extern "C" void clo_dist_set_timing(CloDist* d, cl_bool on) { if (d) d->timing = on != 0; }

extern "C" cl_uint clo_dist_get_phases(CloDist* d, float* out_ms, cl_uint cap) {
	if (!d || !d->ev_valid || !out_ms) return 0;
	cudaEventSynchronize(d->ev[DIST_PHASES]);      /* the local sort is still running when the call returns */
	cl_uint n = 0;
	for (int i = 0; i < DIST_PHASES && n < cap; ++i, ++n) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, d->ev[i], d->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); break; }
		out_ms[n] = ms;
	}
	return n;
}

extern "C" cl_bool clo_dist_get_counts(CloDist* d, cl_ulong* sent, cl_ulong* received) {
	if (!d) return CL_FALSE;
	for (u32 i = 0; i < d->comm.world; ++i) { if (sent) sent[i] = d->last_sent[i]; if (received) received[i] = d->last_recv[i]; }
	return CL_TRUE;
}

extern "C" cl_bool clo_dist_sort_with_device_data(CloDist* d, CCLQueue* cq, CCLBuffer* keys_in, CCLBuffer* payload_in,
		size_t numel, cl_ulong gidx0, CCLBuffer* keys_out, CCLBuffer* payload_out, size_t out_capacity,
		size_t* numel_out, GError** err) {
	if (!d || !clo_handle_alive(d) || !d->recv_k || !cq || !keys_in || !keys_out || !numel_out) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort: invalid arguments (clo_dist_sort_setup first)");
		return CL_FALSE;
	}
	const u32 P = d->comm.world, r = d->comm.rank;
	const size_t kb = (size_t) d->kb;
	if (keys_in->size < numel * kb || (d->with_payload != (payload_in != NULL)) || (d->with_payload && (!payload_out || payload_in->size < numel * 4))) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort: buffers do not match the setup");
		return CL_FALSE;
	}
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	cudaStream_t st = cq->stream;
	const bool tm = d->timing;
	d->ev_valid = false;
	int phase = 0;
	auto mark = [&]() { if (tm && phase <= DIST_PHASES) cudaEventRecord(d->ev[phase], st); ++phase; };
	mark();
	if (P == 1) {
		if (numel > out_capacity || keys_out->size < numel * kb) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort: output too small"); return CL_FALSE; }
		if (numel) {
			if (!d->with_payload) {
				if (!clo_sort_with_device_data(d->sorter, cq, NULL, keys_in, keys_out, numel, 0, err)) return CL_FALSE;
			} else {
				cudaMemcpyAsync(keys_out->ptr, keys_in->ptr, numel * kb, cudaMemcpyDeviceToDevice, st);
				cudaMemcpyAsync(payload_out->ptr, payload_in->ptr, numel * 4, cudaMemcpyDeviceToDevice, st);
				if (!clo_sort_pairs_with_device_data(d->sorter, cq, keys_out, payload_out, numel, err)) return CL_FALSE;
			}
		}
		*numel_out = numel;
		d->last_sent[0] = 0; d->last_recv[0] = numel;
		return CL_TRUE;
	}
	const u32 cap = d->cap;
	const size_t row_w = 2 + cap;
	/* 1) samples -> all-gather */
	clo_dist_sample<<<(cap + 255) / 256, 256, 0, st>>>(keys_in->ptr, (u64) numel, cap, d->kb, d->d_row);
	CLO_COUNT_LAUNCH(1);
	const u64 epoch = ++d->epoch;
	u64* ctl = d->ctl ? (u64*) d->ctl->ptr : nullptr;
	const u64* all_rows = d->d_all;
	u32 all_stride = (u32) row_w;
	if (ctl) {
		/* my row -> slot [rank] of every peer's control buffer; then wait for everybody's */
		clo_dist_bcast<<<1, 1024, 0, st>>>(d->d_row, (u32) row_w, d->d_ctl_ptrs, P, r, CTL_ROWS + (size_t) r * CTL_ROW_MAX, CTL_F_ROWS, epoch);
		clo_dist_wait<<<1, 32, 0, st>>>(ctl, P, CTL_F_ROWS, epoch, d->d_err);
		CLO_COUNT_LAUNCH(2);
		all_rows = ctl + CTL_ROWS;
		all_stride = (u32) CTL_ROW_MAX;
	} else if (d->comm.all_gather_dev(d->comm.user, d->d_row, d->d_all, row_w * 8, st) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_sort: all_gather_dev failed"); return CL_FALSE;
	}
	mark();
	/* 2) splitters */
	cudaMemsetAsync(d->b_splk->ptr, 0, 2 * DIST_MAX_WORLD * 8, st);     /* keys and indices are adjacent */
	clo_dist_splitters<<<(P * cap + 31) / 32, 256, 0, st>>>(all_rows, all_stride, P, cap, r, (u64) gidx0, d->kb, d->b_splk->ptr,
		(u64*) d->b_spli->ptr, d->d_info);
	CLO_COUNT_LAUNCH(1);
	mark();
	/* 3) bucket sizes of this rank */
	if (!clo_sort_partition_count_with_device_data(d->sorter, cq, keys_in, numel, 0, d->b_splk, d->b_spli, P, d->b_counts, err)) return CL_FALSE;
	mark();
	/* 4) everybody's sizes -> where my buckets start at every destination */
	if (ctl) {
		clo_dist_bcast<<<1, 32, 0, st>>>((const u64*) d->b_counts->ptr, DIST_MAX_WORLD, d->d_ctl_ptrs, P, r, CTL_SIZES + (size_t) r * DIST_MAX_WORLD, CTL_F_SIZES, epoch);
		clo_dist_wait<<<1, 32, 0, st>>>(ctl, P, CTL_F_SIZES, epoch, d->d_err);
		clo_dist_slots<<<1, 32, 0, st>>>(ctl + CTL_SIZES, DIST_MAX_WORLD, P, r, (u64) d->capacity, (u64*) d->b_first->ptr, (int*) d->b_ok->ptr, d->h_vec_dev);
		CLO_COUNT_LAUNCH(3);
	} else {
		if (d->comm.all_gather_dev(d->comm.user, d->b_counts->ptr, d->d_M, (size_t) P * 8, st) != 0) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_sort: all_gather_dev failed"); return CL_FALSE;
		}
		clo_dist_slots<<<1, 32, 0, st>>>(d->d_M, P, P, r, (u64) d->capacity, (u64*) d->b_first->ptr, (int*) d->b_ok->ptr, d->h_vec_dev);
		CLO_COUNT_LAUNCH(1);
	}
	mark();
	/* 5) scatter into the receive buffers of the destination ranks (a no-op unless everything fits) */
	if (!clo_sort_partition_scatter_with_device_data(d->sorter, cq, keys_in, payload_in, numel, 0, d->b_splk, d->b_spli, P,
			d->b_first, d->b_ptrs_k, d->with_payload ? d->b_ptrs_p : NULL, d->b_ok, err)) return CL_FALSE;
	mark();
	/* 6) every peer's writes have landed */
	int h_err = 0;
	if (ctl) {
		/* the scatter kernel has completed on this stream: fence, tell everybody, wait for everybody */
		clo_dist_bcast<<<1, 32, 0, st>>>(nullptr, 0, d->d_ctl_ptrs, P, r, 0, CTL_F_DONE, epoch);
		clo_dist_wait<<<1, 32, 0, st>>>(ctl, P, CTL_F_DONE, epoch, d->d_err);
		CLO_COUNT_LAUNCH(2);
		cudaMemcpyAsync(&h_err, d->d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
	} else if (d->comm.barrier_dev(d->comm.user, st) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_sort: barrier_dev failed"); return CL_FALSE;
	}
	if (clo_cuda_failed(cudaStreamSynchronize(st), err, "clo_dist_sort")) return CL_FALSE;     /* the one host synchronisation */
	mark();
	if (h_err) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_sort: a peer did not arrive within the time-out (control flags)");
		return CL_FALSE;
	}
	size_t n_recv = 0;
	for (u32 i = 0; i < P; ++i) { d->last_recv[i] = d->h_vec[i]; d->last_sent[i] = d->h_vec[DIST_MAX_WORLD + i]; n_recv += (size_t) d->h_vec[i]; }
	*numel_out = n_recv;
	if (!d->h_vec[2 * DIST_MAX_WORLD]) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort: a rank would receive %llu elements, the receive buffers hold %zu (nothing was moved)",
			(unsigned long long) d->h_vec[2 * DIST_MAX_WORLD + 1], d->capacity);
		return CL_FALSE;
	}
	if (n_recv > out_capacity || keys_out->size < n_recv * kb || (d->with_payload && payload_out->size < n_recv * 4)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_sort: this rank receives %zu elements, the output holds fewer", n_recv);
		return CL_FALSE;
	}
	/* 7) stable local sort of what arrived (chunks in source-rank order, each in its original order) */
	if (n_recv) {
		if (!d->with_payload) {
			if (!clo_sort_with_device_data(d->sorter, cq, NULL, d->recv_k, keys_out, n_recv, 0, err)) return CL_FALSE;
		} else {
			cudaMemcpyAsync(keys_out->ptr, d->recv_k->ptr, n_recv * kb, cudaMemcpyDeviceToDevice, st);
			cudaMemcpyAsync(payload_out->ptr, d->recv_p->ptr, n_recv * 4, cudaMemcpyDeviceToDevice, st);
			if (!clo_sort_pairs_with_device_data(d->sorter, cq, keys_out, payload_out, n_recv, err)) return CL_FALSE;
		}
	}
	mark();
	d->ev_valid = tm;
	return CL_TRUE;
}

extern "C" CCLEvent* clo_dist_scan_with_device_data(CloDist* d, CloScan* scanner, CCLQueue* cq, CCLBuffer* data_in,
		CCLBuffer* data_out, size_t numel, GError** err) {
	if (!d || !clo_handle_alive(d) || !scanner || !cq || !data_in || !data_out) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "clo_dist_scan: invalid arguments");
		return NULL;
	}
	const u32 P = d->comm.world, r = d->comm.rank;
	if (P == 1) return clo_scan_with_device_data(scanner, cq, NULL, data_in, data_out, numel, 0, err);
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	cudaStream_t st = cq->stream;
	const CloType stype = clo_scan_get_sum_type(scanner);
	const size_t ss = clo_type_sizeof(stype);
	CCLBuffer* b_tot = ccl_buffer_new_wrap(d->ctx, d->d_tot, 8, err);
	CCLBuffer* b_carry = ccl_buffer_new_wrap(d->ctx, d->d_carry, 8, err);
	if (!b_tot || !b_carry) return NULL;
	CCLEvent* evt = NULL;
	cudaMemsetAsync(d->d_tot, 0, 8, st);
	if (numel == 0 || clo_scan_reduce_with_device_data(scanner, cq, data_in, b_tot, numel, err)) {
		/* totals travel as 8-byte words whatever the sum type (little endian: the value sits first) */
		const u64* tots = d->d_tots;
		bool gathered = true;
		if (d->ctl) {
			const u64 epoch = ++d->epoch;
			u64* ctl = (u64*) d->ctl->ptr;
			/* a fast rank may already be in its NEXT scan when this one reads the totals: two slots */
			const size_t slot = CTL_TOTALS + (size_t) (epoch & 1) * DIST_MAX_WORLD;
			clo_dist_bcast<<<1, 32, 0, st>>>(d->d_tot, 1, d->d_ctl_ptrs, P, r, slot + r, CTL_F_TOTALS, epoch);
			clo_dist_wait<<<1, 32, 0, st>>>(ctl, P, CTL_F_TOTALS, epoch, d->d_err);
			CLO_COUNT_LAUNCH(2);
			tots = ctl + slot;
		} else if (d->comm.all_gather_dev(d->comm.user, d->d_tot, d->d_tots, 8, st) != 0) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "clo_dist_scan: all_gather_dev failed");
			gathered = false;
		}
		if (gathered) {
			switch (stype) {
			case CLO_CHAR: case CLO_UCHAR: clo_dist_carry<unsigned char, u32><<<1, 32, 0, st>>>((const unsigned char*) tots, r, (unsigned char*) d->d_carry); break;
			case CLO_SHORT: case CLO_USHORT: clo_dist_carry<unsigned short, u32><<<1, 32, 0, st>>>((const unsigned short*) tots, r, (unsigned short*) d->d_carry); break;
			case CLO_INT: case CLO_UINT: clo_dist_carry<u32, u32><<<1, 32, 0, st>>>((const u32*) tots, r, (u32*) d->d_carry); break;
			case CLO_LONG: case CLO_ULONG: clo_dist_carry<u64, u64><<<1, 32, 0, st>>>((const u64*) tots, r, (u64*) d->d_carry); break;
			case CLO_FLOAT: clo_dist_carry<float, double><<<1, 32, 0, st>>>((const float*) tots, r, (float*) d->d_carry); break;
			case CLO_DOUBLE: clo_dist_carry<double, double><<<1, 32, 0, st>>>((const double*) tots, r, (double*) d->d_carry); break;
			default: break;
			}
			CLO_COUNT_LAUNCH(1);
			(void) ss;
			evt = clo_scan_with_device_data_carry(scanner, cq, data_in, data_out, b_carry, numel, err);
		}
	}
	ccl_buffer_destroy(b_tot);
	ccl_buffer_destroy(b_carry);
	return evt;
}

extern "C" void clo_dist_rng_partition(cl_ulong total_streams, cl_uint rank, cl_uint world, cl_ulong* first, cl_ulong* count) {
	if (world == 0) world = 1;
	const cl_ulong base = total_streams / world, rem = total_streams % world;
	if (first) *first = rank * base + (rank < rem ? rank : rem);
	if (count) *count = base + (rank < rem ? 1 : 0);
}
