/*
 * radix_v6.cu -- translation unit of the keys-only two-barrier onesweep kernel
 * (radix_v6.cuh) and its launcher.  Replaces the per-digit {localsort, histogram, scan,
 * scatter} launches of the reference (/root/reference/src/cl_ops/sort/clo_sort_satradix.c:264-313)
 * for sorts whose element is the key.
 */
#include "clo_internal.h"
#include "device_utils.cuh"
#include "sort_common.h"

#include <cstdlib>

using namespace clo;
#ifndef CLO_IPT_U32
#define CLO_IPT_U32 16
#endif

namespace {
#include "radix_prop.cuh"
#include "radix_v6.cuh"

/* 1 when this device serves same-address shared atomics of a warp instruction in lane order
 * (checked once per device and process), 0 when not or when the check could not run */
int atomic_order_selftest(cudaStream_t stream) {
	static int result[64];
	static bool done[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	if (done[dev]) return result[dev];
	u32* d = nullptr;
	u32 h[2] = { 1, 0 };
	int r = 0;
	const int blocks = 64, threads = 512, iters = 256;
	if (cudaMalloc(&d, 8) == cudaSuccess) {
		if (cudaMemsetAsync(d, 0, 8, stream) == cudaSuccess) {
			clo_radix_atomic_order_selftest_kernel<<<blocks, threads, 0, stream>>>(d, 20261018u, iters);
			if (cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
				cudaStreamSynchronize(stream) == cudaSuccess)
				r = (h[0] == 0 && h[1] == (u32) (blocks * threads * iters)) ? 1 : 0;
		}
		cudaFree(d);
	}
	result[dev] = r;
	done[dev] = true;
	return r;
}

template <typename ElemT, typename LbT, int THREADS, int IPT, bool HAS_VAL, bool VERIFY>
cudaError_t launch_v6(int tile, const ElemT* in, ElemT* out, const u32* vin, u32* vout, size_t n, LbT* agg, LbT* pref, u32* ticket,
		const u64* bins, u32 start_bit, u32 dmask, int* err, int sm_count, int prof_on, int flags, cudaStream_t stream,
		const void* chain_in, void* chain_out, void* out_alt, u32* vout_alt) {
	if (tile != THREADS * IPT) return cudaErrorInvalidValue;
	constexpr size_t SMEM = onesweep_v6_smem<ElemT, THREADS, IPT, LbT, HAS_VAL>();
	auto kern = clo_radix_onesweep_v6<ElemT, LbT, THREADS, IPT, HAS_VAL, VERIFY>;
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
	constexpr int V6_NUM_PROP = v6_num_prop(THREADS, (int) sizeof(LbT));
	workers = workers > (size_t) V6_NUM_PROP ? workers - V6_NUM_PROP : 1;
	if (workers > tiles) workers = tiles;
	kern<<<(unsigned) (V6_NUM_PROP + workers), THREADS, SMEM, stream>>>(in, out, vin, vout, n, (u32) tiles,
		agg, pref, ticket, bins, start_bit, dmask, err, prof_on, flags,
		(const V6Chain*) chain_in, (V6Chain*) chain_out, (ElemT*) out_alt, vout_alt);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}
} // namespace

cudaError_t clo_radix_v6_chain_fixup(const void* chain, void* dst, void* vdst, size_t key_bytes, size_t val_bytes,
		int sm_count, cudaStream_t stream) {
	clo_radix_chain_fixup<<<(unsigned) sm_count * 4, 512, 0, stream>>>((const V6Chain*) chain, dst, vdst, key_bytes, val_bytes);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

/* tile sizes: keys only 512 x 16 (4-byte keys) / 512 x 8 (8-byte keys); with a u32 payload
 * 512 x 8 / 512 x 6, so that two CTAs and their double staging buffers fit one SM */
cudaError_t clo_radix_v6_pass(int elem_size, int wide, int tile, const void* in, void* out,
		const uint32_t* vin, uint32_t* vout, size_t n,
		void* agg, void* pref, uint32_t* ticket, const unsigned long long* bins, uint32_t start_bit,
		uint32_t dmask, int* err, int sm_count, int prof_on, int flags, cudaStream_t stream,
		const void* chain_in, void* chain_out, void* out_alt, uint32_t* vout_alt) {
	/* the per-tile stability check is compiled out when the device passed the atomic-order
	 * self-test; CLO_RADIX_VERIFY=1 keeps it, and so do the fault hooks of the tests */
	static int force_verify = -1;
	if (force_verify < 0) { const char* e = getenv("CLO_RADIX_VERIFY"); force_verify = (e && *e == '1') ? 1 : 0; }
	const bool verify = force_verify || (flags & (8 | 16)) || atomic_order_selftest(stream) != 1;
#define CLO_V6_GO(ET, LT, IPT_, HV) (verify \
	? launch_v6<ET, LT, 512, IPT_, HV, true>(tile, (const ET*) in, (ET*) out, vin, vout, n, (LT*) agg, (LT*) pref, \
		ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream, chain_in, chain_out, out_alt, vout_alt) \
	: launch_v6<ET, LT, 512, IPT_, HV, false>(tile, (const ET*) in, (ET*) out, vin, vout, n, (LT*) agg, (LT*) pref, \
		ticket, bins, start_bit, dmask, err, sm_count, prof_on, flags, stream, chain_in, chain_out, out_alt, vout_alt))
	if (elem_size == 4 && !vin) return wide ? CLO_V6_GO(u32, u64, CLO_IPT_U32, false) : CLO_V6_GO(u32, u32, CLO_IPT_U32, false);
	if (elem_size == 8 && !vin) return wide ? CLO_V6_GO(u64, u64, 10, false) : CLO_V6_GO(u64, u32, 10, false);
	if (elem_size == 4 && vin) return wide ? CLO_V6_GO(u32, u64, 8, true) : CLO_V6_GO(u32, u32, 8, true);
	if (elem_size == 8 && vin) return wide ? CLO_V6_GO(u64, u64, 6, true) : CLO_V6_GO(u64, u32, 6, true);
#undef CLO_V6_GO
	return cudaErrorInvalidValue;
}
