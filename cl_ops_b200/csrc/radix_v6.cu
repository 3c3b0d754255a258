/*
 * radix_v6.cu -- translation unit of the keys-only two-barrier onesweep kernel
 * (radix_v6.cuh) and its launcher.  Replaces the per-digit {localsort, histogram, scan,
 * scatter} launches of the reference (/root/reference/src/cl_ops/sort/clo_sort_satradix.c:264-313)
 * for sorts whose element is the key.
 */
#include "clo_internal.h"
#include "device_utils.cuh"
#include "sort_common.h"

using namespace clo;

namespace {
#include "radix_prop.cuh"
#include "radix_v6.cuh"

template <typename ElemT, typename LbT, int THREADS, int IPT, int ABL = 0>
cudaError_t launch_v6(int tile, const ElemT* in, ElemT* out, size_t n, LbT* agg, LbT* pref, u32* ticket,
		const u64* bins, u32 start_bit, u32 dmask, int* err, int sm_count, int prof_on, int flags, cudaStream_t stream) {
	if (tile != THREADS * IPT) return cudaErrorInvalidValue;
	constexpr size_t SMEM = onesweep_v6_smem<ElemT, THREADS, IPT, LbT>();
	auto kern = clo_radix_onesweep_v6<ElemT, LbT, THREADS, IPT, ABL>;
	static bool configured[64] = {};
	static int ctas_per_sm[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	if (!configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM);
		if (e != cudaSuccess) return e;
		int k = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, THREADS, SMEM) != cudaSuccess || k < 1) k = 1;
		ctas_per_sm[dev] = k;
		configured[dev] = true;
	}
	const size_t tiles = (n + (size_t) tile - 1) / (size_t) tile;
	size_t workers = (size_t) sm_count * ctas_per_sm[dev];
	constexpr int V6_NUM_PROP = v6_num_prop(THREADS);
	workers = workers > (size_t) V6_NUM_PROP ? workers - V6_NUM_PROP : 1;
	if (workers > tiles) workers = tiles;
	kern<<<(unsigned) (V6_NUM_PROP + workers), THREADS, SMEM, stream>>>(in, out, n, (u32) tiles,
		agg, pref, ticket, bins, start_bit, dmask, err, prof_on, flags);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}
} // namespace

cudaError_t clo_radix_v6_pass(int elem_size, int wide, int tile, const void* in, void* out, size_t n,
		void* agg, void* pref, uint32_t* ticket, const unsigned long long* bins, uint32_t start_bit,
		uint32_t dmask, int* err, int sm_count, int prof_on, int flags, cudaStream_t stream) {
#ifdef CLO_V6_ABLATION
	if (elem_size == 4 && !wide && (flags >> 8)) {
		switch (flags >> 8) {
		case 1: return launch_v6<u32, u32, 512, 16, 1>(tile, (const u32*) in, (u32*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		case 2: return launch_v6<u32, u32, 512, 16, 2>(tile, (const u32*) in, (u32*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		case 3: return launch_v6<u32, u32, 512, 16, 3>(tile, (const u32*) in, (u32*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		case 4: return launch_v6<u32, u32, 512, 16, 4>(tile, (const u32*) in, (u32*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		}
	}
#endif
	if (elem_size == 4) {
		if (wide) return launch_v6<u32, u64, 512, 16>(tile, (const u32*) in, (u32*) out, n, (u64*) agg, (u64*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		return launch_v6<u32, u32, 512, 16>(tile, (const u32*) in, (u32*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
	}
	if (elem_size == 8) {
		if (wide) return launch_v6<u64, u64, 512, 8>(tile, (const u64*) in, (u64*) out, n, (u64*) agg, (u64*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
		return launch_v6<u64, u32, 512, 8>(tile, (const u64*) in, (u64*) out, n, (u32*) agg, (u32*) pref, ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream);
	}
	return cudaErrorInvalidValue;
}
