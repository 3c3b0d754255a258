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
 * Here: all steps with stride < TILE run in shared memory (one launch sorts
 * every TILE-sized block through stage log2(TILE); one launch finishes each
 * later stage), the remaining steps are one global compare-exchange launch each.
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
};

CloBitonicState* clo_bitonic_state_new() { return new CloBitonicState(); }

void clo_bitonic_state_free(CloBitonicState* st) {
	if (!st) return;
	st->padded.release();
	delete st;
}

namespace {
template <typename ElemT>
cudaError_t bitonic_typed(CloBitonicState* st, const CloKeySpec& ks, void* data, size_t n, cudaStream_t stream) {
	size_t np2 = 1;
	while (np2 < n) np2 <<= 1;
	if (np2 == n) return bitonic_run<ElemT, false>((ElemT*) data, nullptr, np2, ks, stream);
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

cudaError_t clo_bitonic_sort(CloBitonicState* st, size_t elem_size, const CloKeySpec& ks,
		void* data, size_t n, cudaStream_t stream) {
	if (n < 2) return cudaSuccess;
	switch (elem_size) {
	case 1: return bitonic_typed<unsigned char>(st, ks, data, n, stream);
	case 2: return bitonic_typed<unsigned short>(st, ks, data, n, stream);
	case 4: return bitonic_typed<u32>(st, ks, data, n, stream);
	case 8: return bitonic_typed<u64>(st, ks, data, n, stream);
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
