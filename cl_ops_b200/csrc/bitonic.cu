/*
 * bitonic.cu -- the canonical bitonic network of the reference's "sbitonic" and
 * "abitonic" sorters, and the "gselect" stable rank sort.
 *
 * Network (identical compare-exchanges, hence identical results for any
 * compare / get_key): /root/reference/src/cl_ops/sort/clo_sort_sbitonic.cl:38-69,
 * host loop clo_sort_sbitonic.c:73-118; the abitonic kernels
 * (clo_sort_abitonic.cl:31-1067) fuse steps of the same network.
 *   for stage s = 1..log2(N), step t = s..1, pair p in [0, N/2):
 *     stride = 2^(t-1); i1 = p + (p / stride) * stride; i2 = i1 + stride
 *     desc = (p >> (s-1)) & 1   ( == (i1 >> s) & 1 )
 *     swap iff COMPARE(key(e[i1]), key(e[i2])) XOR desc
 *
 * Power-of-two N (the only N the reference defines): ONE persistent cooperative launch
 * (clo_bitonic_fused).  A thread keeps 2^k elements in registers and runs k consecutive steps
 * on them without touching memory -- the reference's own idea for its abit_priv_2s4v / 3s8v /
 * 4s16v kernels (clo_sort_abitonic.cl:607-650) -- with k <= 4 (abitonic's maxps option caps
 * it).  Steps with stride < TILE exchange through a padded shared-memory tile (one CTA barrier
 * per k steps), larger strides through global memory (L2 resident up to ~100 MB) with a grid
 * barrier per k steps.  2^20 keys: 1 launch and 17 grid barriers instead of 45 launches (and
 * the reference's 210).
 * Other N: the network on an array padded with "+infinity" flags, one launch per global step
 * (clo_bitonic_local / clo_bitonic_global), as before.
 *
 * gselect: clo_sort_gselect.cl:38-57.
 */
#include "clo_internal.h"
#include "device_utils.cuh"
#include "sort_common.h"

using namespace clo;

namespace {

const int BT_THREADS = 512;
const int BT_LOG_TILE = 12;
const int BT_TILE = 1 << BT_LOG_TILE;

/* "a must come after b": CLO_SORT_COMPARE(key(a), key(b)).  Padding elements
 * (only present when N is not a power of two) come after everything. */
template <typename ElemT, bool PADDED>
__device__ __forceinline__ bool must_swap(ElemT a, ElemT b, unsigned char pa, unsigned char pb, const CloKeySpec& ks) {
	const u64 ka = clo_ordered_key(clo_extract_key((u64) a, ks), ks);
	const u64 kb = clo_ordered_key(clo_extract_key((u64) b, ks), ks);
	if (PADDED) return (pa > pb) || (pa == pb && ka > kb);
	return ka > kb;
}

/* FULL: stages 1..last_stage, all steps (block sort).  !FULL: stage `last_stage`,
 * steps min(last_stage, LOG_TILE)..1 (finish one stage). */
template <typename ElemT, bool PADDED, bool FULL>
__global__ void __launch_bounds__(BT_THREADS)
clo_bitonic_local(ElemT* __restrict__ data, unsigned char* __restrict__ pad, size_t np2,
		int last_stage, CloKeySpec ks) {
	__shared__ ElemT s[BT_TILE];
	__shared__ unsigned char sp[PADDED ? BT_TILE : 1];
	const size_t base = (size_t) blockIdx.x * BT_TILE;
	const int cnt = (np2 - base) < (size_t) BT_TILE ? (int) (np2 - base) : BT_TILE;
	for (int i = threadIdx.x; i < cnt; i += BT_THREADS) {
		s[i] = data[base + i];
		if (PADDED) sp[i] = pad[base + i];
	}
	__syncthreads();
	const int first_stage = FULL ? 1 : last_stage;
	for (int stage = first_stage; stage <= last_stage; ++stage) {
		const int top = (FULL || stage < BT_LOG_TILE) ? (stage < BT_LOG_TILE ? stage : BT_LOG_TILE) : BT_LOG_TILE;
		for (int step = top; step > 0; --step) {
			const int stride = 1 << (step - 1);
			for (int p = threadIdx.x; p < cnt / 2; p += BT_THREADS) {
				const int i1 = p + (p / stride) * stride;
				const int i2 = i1 + stride;
				const bool desc = ((base + (size_t) i1) >> stage) & 1;
				const ElemT a = s[i1], b = s[i2];
				const unsigned char pa = PADDED ? sp[i1] : 0, pb = PADDED ? sp[i2] : 0;
				if (must_swap<ElemT, PADDED>(a, b, pa, pb, ks) != desc) {
					s[i1] = b; s[i2] = a;
					if (PADDED) { sp[i1] = pb; sp[i2] = pa; }
				}
			}
			__syncthreads();
		}
	}
	for (int i = threadIdx.x; i < cnt; i += BT_THREADS) {
		data[base + i] = s[i];
		if (PADDED) pad[base + i] = sp[i];
	}
}

/* one compare-exchange step with stride >= TILE */
template <typename ElemT, bool PADDED>
__global__ void __launch_bounds__(256)
clo_bitonic_global(ElemT* __restrict__ data, unsigned char* __restrict__ pad, size_t npairs,
		int stage, int step, CloKeySpec ks) {
	const size_t p = (size_t) blockIdx.x * 256 + threadIdx.x;
	if (p >= npairs) return;
	const size_t stride = (size_t) 1 << (step - 1);
	const size_t i1 = p + (p / stride) * stride;
	const size_t i2 = i1 + stride;
	const bool desc = (p >> (stage - 1)) & 1;
	const ElemT a = data[i1], b = data[i2];
	const unsigned char pa = PADDED ? pad[i1] : 0, pb = PADDED ? pad[i2] : 0;
	if (must_swap<ElemT, PADDED>(a, b, pa, pb, ks) != desc) {
		data[i1] = b; data[i2] = a;
		if (PADDED) { pad[i1] = pb; pad[i2] = pa; }
	}
}

/* ------------------------------------------------------------------ fused persistent network */

/* a must come after b -- specialised when the key is the element itself, unsigned, ascending */
template <typename ElemT, bool SIMPLE>
__device__ __forceinline__ bool bf_after(ElemT a, ElemT b, const CloKeySpec& ks) {
	if (SIMPLE) return a > b;
	return clo_ordered_key(clo_extract_key((u64) a, ks), ks) > clo_ordered_key(clo_extract_key((u64) b, ks), ks);
}

/* k steps of one stage on 2^k register-resident elements that are `spacing` apart in the array:
 * register strides 2^(k-1) .. 1 are the array strides of steps top .. top-k+1.  The direction
 * bit (index >> stage) is the same for the whole group because stage >= top. */
template <typename ElemT, bool SIMPLE, int K>
__device__ __forceinline__ void bf_steps(ElemT (&v)[1 << K], bool desc, const CloKeySpec& ks) {
#pragma unroll
	for (int r = 1 << (K - 1); r > 0; r >>= 1) {
#pragma unroll
		for (int j = 0; j < (1 << K); ++j) {
			if ((j & r) == 0) {
				const ElemT a = v[j], b = v[j + r];
				if (bf_after<ElemT, SIMPLE>(a, b, ks) != desc) { v[j] = b; v[j + r] = a; }
			}
		}
	}
}

/* shared-memory tile with one padding word per 32: power-of-two strides stay (almost) conflict free */
__device__ __forceinline__ u32 bf_pad(u32 i) { return i + (i >> 5); }

/* steps top .. top-K+1 of `stage` on a tile in shared memory; gbase = array index of the tile's element 0 */
template <typename ElemT, bool SIMPLE, int K, int THREADS>
__device__ __forceinline__ void bf_tile_pass(ElemT* s, u32 tile_elems, size_t gbase, int stage, int top, const CloKeySpec& ks) {
	const int low_bits = top - K;                                   /* spacing = 2^low_bits */
	const u32 groups = tile_elems >> K;
	for (u32 g = threadIdx.x; g < groups; g += THREADS) {
		const u32 base = ((g >> low_bits) << top) | (g & ((1u << low_bits) - 1u));
		ElemT v[1 << K];
#pragma unroll
		for (int j = 0; j < (1 << K); ++j) v[j] = s[bf_pad(base + ((u32) j << low_bits))];
		bf_steps<ElemT, SIMPLE, K>(v, ((gbase + base) >> stage) & 1, ks);
#pragma unroll
		for (int j = 0; j < (1 << K); ++j) s[bf_pad(base + ((u32) j << low_bits))] = v[j];
	}
}

/* steps top .. 1 of `stage` on the tile, in passes of at most KMAX steps */
template <typename ElemT, bool SIMPLE, int THREADS>
__device__ __forceinline__ void bf_tile_steps(ElemT* s, u32 tile_elems, size_t gbase, int stage, int top, int kmax, const CloKeySpec& ks) {
	while (top > 0) {
		/* 13 = 5 + 4 + 4 and 9 = 5 + 4: one pass fewer than with 4 steps each (kmax 5 = default) */
		int k = top < kmax ? top : kmax;
		if (kmax == 5 && k == 5 && (top % 4) != 1) k = 4;
		switch (k) {
		case 5: bf_tile_pass<ElemT, SIMPLE, 5, THREADS>(s, tile_elems, gbase, stage, top, ks); break;
		case 4: bf_tile_pass<ElemT, SIMPLE, 4, THREADS>(s, tile_elems, gbase, stage, top, ks); break;
		case 3: bf_tile_pass<ElemT, SIMPLE, 3, THREADS>(s, tile_elems, gbase, stage, top, ks); break;
		case 2: bf_tile_pass<ElemT, SIMPLE, 2, THREADS>(s, tile_elems, gbase, stage, top, ks); break;
		default: bf_tile_pass<ElemT, SIMPLE, 1, THREADS>(s, tile_elems, gbase, stage, top, ks); break;
		}
		top -= k;
		__syncthreads();
	}
}

/* steps top .. top-K+1 of `stage` straight on global memory (strides >= TILE) */
template <typename ElemT, bool SIMPLE, int K, int THREADS>
__device__ __forceinline__ void bf_global_pass(ElemT* __restrict__ data, size_t n, int stage, int top, const CloKeySpec& ks) {
	const int low_bits = top - K;
	const size_t groups = n >> K;
	for (size_t g = (size_t) blockIdx.x * THREADS + threadIdx.x; g < groups; g += (size_t) gridDim.x * THREADS) {
		const size_t base = ((g >> low_bits) << top) | (g & (((size_t) 1 << low_bits) - 1));
		ElemT v[1 << K];
#pragma unroll
		for (int j = 0; j < (1 << K); ++j) v[j] = data[base + ((size_t) j << low_bits)];
		bf_steps<ElemT, SIMPLE, K>(v, (base >> stage) & 1, ks);
#pragma unroll
		for (int j = 0; j < (1 << K); ++j) data[base + ((size_t) j << low_bits)] = v[j];
	}
}

/* all CTAs of the (cooperative, fully resident) grid meet; `target` = arrivals expected in total */
__device__ __forceinline__ void bf_grid_barrier(unsigned long long* counter, unsigned long long target) {
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(counter, 1ull);
		while (*((volatile unsigned long long*) counter) < target) { }
		__threadfence();
	}
	__syncthreads();
}

template <typename ElemT, bool SIMPLE, int THREADS, int LOG_TILE>
__global__ void __launch_bounds__(THREADS)
clo_bitonic_fused(ElemT* __restrict__ data, size_t n, int log_n, int kmax, CloKeySpec ks,
		unsigned long long* __restrict__ bar_counter, unsigned long long bar_start) {
	extern __shared__ __align__(16) unsigned char bf_smem[];
	ElemT* s = reinterpret_cast<ElemT*>(bf_smem);
	constexpr u32 TILE = 1u << LOG_TILE;
	const u32 tile_elems = n < (size_t) TILE ? (u32) n : TILE;
	const int log_tile = log_n < LOG_TILE ? log_n : LOG_TILE;
	const size_t tiles = n / tile_elems;
	unsigned long long arrivals = bar_start;
	auto load_tile = [&](size_t t) {
		for (u32 i = threadIdx.x; i < tile_elems; i += THREADS) s[bf_pad(i)] = data[t * tile_elems + i];
		__syncthreads();
	};
	auto store_tile = [&](size_t t) {
		for (u32 i = threadIdx.x; i < tile_elems; i += THREADS) data[t * tile_elems + i] = s[bf_pad(i)];
		__syncthreads();
	};
	/* stages 1 .. log_tile: every tile on its own */
	for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
		load_tile(t);
		for (int stage = 1; stage <= log_tile; ++stage)
			bf_tile_steps<ElemT, SIMPLE, THREADS>(s, tile_elems, t * tile_elems, stage, stage, kmax, ks);
		store_tile(t);
	}
	/* later stages: global steps in register-fused passes, then the tile-local rest of the stage */
	for (int stage = log_tile + 1; stage <= log_n; ++stage) {
		int top = stage;
		while (top > log_tile) {
			arrivals += gridDim.x;
			bf_grid_barrier(bar_counter, arrivals);
			const int k = (top - log_tile) < kmax ? (top - log_tile) : kmax;
			switch (k) {
			case 5: bf_global_pass<ElemT, SIMPLE, 5, THREADS>(data, n, stage, top, ks); break;
			case 4: bf_global_pass<ElemT, SIMPLE, 4, THREADS>(data, n, stage, top, ks); break;
			case 3: bf_global_pass<ElemT, SIMPLE, 3, THREADS>(data, n, stage, top, ks); break;
			case 2: bf_global_pass<ElemT, SIMPLE, 2, THREADS>(data, n, stage, top, ks); break;
			default: bf_global_pass<ElemT, SIMPLE, 1, THREADS>(data, n, stage, top, ks); break;
			}
			top -= k;
		}
		arrivals += gridDim.x;
		bf_grid_barrier(bar_counter, arrivals);
		for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
			load_tile(t);
			bf_tile_steps<ElemT, SIMPLE, THREADS>(s, tile_elems, t * tile_elems, stage, log_tile, kmax, ks);
			store_tile(t);
		}
	}
}

/* grid barriers one launch of clo_bitonic_fused goes through */
static int bf_barriers(int log_n, int log_tile, int kmax) {
	int b = 0;
	for (int stage = log_tile + 1; stage <= log_n; ++stage) b += (stage - log_tile + kmax - 1) / kmax + 1;
	return b;
}

__global__ void clo_bitonic_init_pad(unsigned char* __restrict__ pad, size_t n, size_t np2) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < np2) pad[i] = i >= n ? 1 : 0;
}

/* rank = #{ i : COMPARE(key_gid, key_i) || (key_i == key_gid && i < gid) } */
template <typename ElemT>
__global__ void __launch_bounds__(256)
clo_gselect_kernel(const ElemT* __restrict__ in, ElemT* __restrict__ out, size_t n, CloKeySpec ks) {
	__shared__ u64 s_key[256];
	const size_t gid = (size_t) blockIdx.x * 256 + threadIdx.x;
	ElemT mine = ElemT(0);
	u64 kg = 0;
	if (gid < n) {
		mine = in[gid];
		kg = clo_ordered_key(clo_extract_key((u64) mine, ks), ks);
	}
	size_t rank = 0;
	for (size_t t0 = 0; t0 < n; t0 += 256) {
		const size_t i = t0 + threadIdx.x;
		__syncthreads();
		if (i < n) s_key[threadIdx.x] = clo_ordered_key(clo_extract_key((u64) in[i], ks), ks);
		__syncthreads();
		const int lim = (n - t0) < 256 ? (int) (n - t0) : 256;
		for (int j = 0; j < lim; ++j) {
			const u64 ki = s_key[j];
			if (kg > ki || (ki == kg && (t0 + j) < gid)) ++rank;
		}
	}
	if (gid < n) out[rank] = mine;
}

template <typename ElemT, bool PADDED>
cudaError_t bitonic_run(ElemT* d, unsigned char* pad, size_t np2, const CloKeySpec& ks, cudaStream_t stream) {
	int log = 0;
	while (((size_t) 1 << log) < np2) ++log;
	const unsigned blocks = (unsigned) ((np2 + BT_TILE - 1) / BT_TILE);
	const int first = log < BT_LOG_TILE ? log : BT_LOG_TILE;
	unsigned long long launches = 0;
	if (first >= 1) {
		clo_bitonic_local<ElemT, PADDED, true><<<blocks, BT_THREADS, 0, stream>>>(d, pad, np2, first, ks);
		++launches;
	}
	const size_t npairs = np2 / 2;
	for (int stage = BT_LOG_TILE + 1; stage <= log; ++stage) {
		for (int step = stage; step > BT_LOG_TILE; --step) {
			clo_bitonic_global<ElemT, PADDED><<<(unsigned) ((npairs + 255) / 256), 256, 0, stream>>>(d, pad, npairs, stage, step, ks);
			++launches;
		}
		clo_bitonic_local<ElemT, PADDED, false><<<blocks, BT_THREADS, 0, stream>>>(d, pad, np2, stage, ks);
		++launches;
	}
	CLO_COUNT_LAUNCH(launches);
	return cudaGetLastError();
}

} // namespace

struct CloBitonicState {
	CloScratch padded;   /* np2 elements + np2 pad flags, only for non power-of-two n */
	CloScratch barrier;  /* arrival counter of the fused kernel's grid barrier (monotonic) */
	unsigned long long arrivals = 0;   /* value the counter has when the next launch starts */
	int kmax = 5;        /* steps fused in registers (abitonic option maxps: 1..4; 5 = the library's default) */
	int max_log_tile = 13;   /* cap of the shared-memory tile, log2 elements (abitonic option maxsfs) */
};

#ifndef CLO_BITONIC_ONLY
CloBitonicState* clo_bitonic_state_new() { return new CloBitonicState(); }

void clo_bitonic_state_free(CloBitonicState* st) {
	if (!st) return;
	st->padded.release();
	st->barrier.release();
	delete st;
}

void clo_bitonic_set_fusion(CloBitonicState* st, int max_private_steps, int max_local_steps) {
	if (!st) return;
	st->kmax = max_private_steps < 1 ? 1 : (max_private_steps > 5 ? 5 : max_private_steps);
	st->max_log_tile = max_local_steps < 5 ? 5 : (max_local_steps > 13 ? 13 : max_local_steps);
}

#endif

namespace {
/* power-of-two n: the whole network as one cooperative launch; false when that is not possible */
template <typename ElemT, bool SIMPLE, int LOG_TILE>
bool bitonic_fused_launch(CloBitonicState* st, const CloKeySpec& ks, ElemT* data, size_t n, cudaStream_t stream, cudaError_t& rc) {
	constexpr int THREADS = 512;
	constexpr size_t SMEM = ((size_t) (1u << LOG_TILE) + ((size_t) (1u << LOG_TILE) >> 5) + 1) * sizeof(ElemT);
	auto kern = clo_bitonic_fused<ElemT, SIMPLE, THREADS, LOG_TILE>;
	static bool configured[64] = {};
	static int resident[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	if (!configured[dev]) {
		int coop = 0, sms = 0, k = 0;
		cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM) != cudaSuccess ||
			cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, THREADS, SMEM) != cudaSuccess) k = 0;
		resident[dev] = coop ? k * sms : 0;
		configured[dev] = true;
	}
	if (resident[dev] < 1) return false;
	int log_n = 0;
	while (((size_t) 1 << log_n) < n) ++log_n;
	if (!st->barrier.ptr) {
		if ((rc = st->barrier.reserve(64)) != cudaSuccess) return true;
		if ((rc = cudaMemsetAsync(st->barrier.ptr, 0, 64, stream)) != cudaSuccess) return true;
		st->arrivals = 0;
	}
	const size_t tile = n < ((size_t) 1 << LOG_TILE) ? n : ((size_t) 1 << LOG_TILE);
	const size_t tiles = n / tile;
	const unsigned grid = (unsigned) (tiles < (size_t) resident[dev] ? tiles : (size_t) resident[dev]);
	int kmax = st->kmax;
	unsigned long long* counter = (unsigned long long*) st->barrier.ptr;
	unsigned long long start = st->arrivals;
	CloKeySpec k = ks;
	void* args[] = { &data, &n, &log_n, &kmax, &k, &counter, &start };
	rc = cudaLaunchCooperativeKernel((const void*) kern, dim3(grid), dim3(THREADS), args, SMEM, stream);
	if (rc == cudaSuccess) {
		st->arrivals += (unsigned long long) grid * (unsigned long long) bf_barriers(log_n, log_n < LOG_TILE ? log_n : LOG_TILE, kmax);
		CLO_COUNT_LAUNCH(1);
	}
	return true;
}

template <typename ElemT>
bool bitonic_fused(CloBitonicState* st, const CloKeySpec& ks, ElemT* data, size_t n, cudaStream_t stream, cudaError_t& rc) {
	constexpr int FULL_LOG = sizeof(ElemT) == 8 ? 12 : 13;          /* 32 KB tile */
	const bool simple = ks.identity && ks.key_kind == CLO_KIND_UNSIGNED && !ks.descending;
	const int lt = st->max_log_tile < FULL_LOG ? st->max_log_tile : FULL_LOG;
	if (lt >= FULL_LOG)
		return simple ? bitonic_fused_launch<ElemT, true, FULL_LOG>(st, ks, data, n, stream, rc)
			: bitonic_fused_launch<ElemT, false, FULL_LOG>(st, ks, data, n, stream, rc);
	/* abitonic maxsfs below the full tile: one smaller tile shape */
	return simple ? bitonic_fused_launch<ElemT, true, 9>(st, ks, data, n, stream, rc)
		: bitonic_fused_launch<ElemT, false, 9>(st, ks, data, n, stream, rc);
}

template <typename ElemT>
cudaError_t bitonic_typed(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream) {
	size_t np2 = 1;
	while (np2 < n) np2 <<= 1;
	if (np2 == n) {
		cudaError_t rc = cudaSuccess;
		if (bitonic_fused<ElemT>(st, ks, (ElemT*) data, n, stream, rc)) return rc;
		return bitonic_run<ElemT, false>((ElemT*) data, nullptr, np2, ks, stream);
	}
	cudaError_t e;
	if ((e = st->padded.reserve(np2 * sizeof(ElemT) + np2)) != cudaSuccess) return e;
	ElemT* tmp = (ElemT*) st->padded.ptr;
	unsigned char* pad = (unsigned char*) st->padded.ptr + np2 * sizeof(ElemT);
	if ((e = cudaMemcpyAsync(tmp, data, n * sizeof(ElemT), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
	if ((e = cudaMemsetAsync(tmp + n, 0, (np2 - n) * sizeof(ElemT), stream)) != cudaSuccess) return e;
	clo_bitonic_init_pad<<<(unsigned) ((np2 + 255) / 256), 256, 0, stream>>>(pad, n, np2);
	CLO_COUNT_LAUNCH(1);
	if ((e = bitonic_run<ElemT, true>(tmp, pad, np2, ks, stream)) != cudaSuccess) return e;
	return cudaMemcpyAsync(data, tmp, n * sizeof(ElemT), cudaMemcpyDeviceToDevice, stream);
}
}

/* The fused kernel is heavy to compile (k <= 5 unrolled register networks per element width and
 * comparator form: 7 minutes in one translation unit), so the Makefile compiles this file once
 * per element width with -DCLO_BITONIC_ONLY=<bytes>, in parallel; each of those objects holds one
 * clo_bitonic_typed_<bytes>, and the plain object the entry points that dispatch to them. */
#ifdef CLO_BITONIC_ONLY
#define CLO_BT_CAT2(a, b) a##b
#define CLO_BT_CAT(a, b) CLO_BT_CAT2(a, b)
#if CLO_BITONIC_ONLY == 1
typedef unsigned char CloBitonicOnlyT;
#elif CLO_BITONIC_ONLY == 2
typedef unsigned short CloBitonicOnlyT;
#elif CLO_BITONIC_ONLY == 4
typedef u32 CloBitonicOnlyT;
#else
typedef u64 CloBitonicOnlyT;
#endif
cudaError_t CLO_BT_CAT(clo_bitonic_typed_, CLO_BITONIC_ONLY)(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream) {
	return bitonic_typed<CloBitonicOnlyT>(st, ks, data, n, stream);
}
#else
cudaError_t clo_bitonic_typed_1(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream);
cudaError_t clo_bitonic_typed_2(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream);
cudaError_t clo_bitonic_typed_4(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream);
cudaError_t clo_bitonic_typed_8(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream);

cudaError_t clo_bitonic_sort(CloBitonicState* st, size_t elem_size, const CloKeySpec& ks,
		void* data, size_t n, cudaStream_t stream) {
	if (n < 2) return cudaSuccess;
	switch (elem_size) {
	case 1: return clo_bitonic_typed_1(st, ks, data, n, stream);
	case 2: return clo_bitonic_typed_2(st, ks, data, n, stream);
	case 4: return clo_bitonic_typed_4(st, ks, data, n, stream);
	case 8: return clo_bitonic_typed_8(st, ks, data, n, stream);
	default: return cudaErrorInvalidValue;
	}
}

cudaError_t clo_gselect_sort(size_t elem_size, const CloKeySpec& ks, const void* in, void* out,
		size_t n, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	const unsigned blocks = (unsigned) ((n + 255) / 256);
	switch (elem_size) {
	case 1: clo_gselect_kernel<unsigned char><<<blocks, 256, 0, stream>>>((const unsigned char*) in, (unsigned char*) out, n, ks); break;
	case 2: clo_gselect_kernel<unsigned short><<<blocks, 256, 0, stream>>>((const unsigned short*) in, (unsigned short*) out, n, ks); break;
	case 4: clo_gselect_kernel<u32><<<blocks, 256, 0, stream>>>((const u32*) in, (u32*) out, n, ks); break;
	case 8: clo_gselect_kernel<u64><<<blocks, 256, 0, stream>>>((const u64*) in, (u64*) out, n, ks); break;
	default: return cudaErrorInvalidValue;
	}
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}
#endif
