/*
 * scan.cu -- clo_scan: exclusive prefix sum as ONE single-pass decoupled
 * look-back kernel (replaces the reference's three Blelloch kernels,
 * /root/reference/src/cl_ops/scan/clo_scan_blelloch.cl:49-211 and the host loop
 * clo_scan_blelloch.c:78-214; API object: clo_scan_abstract.c:74-567).
 *
 * Semantics kept: out[0] = 0, out[i] = sum_{j<i} (SUM)in[j], SUM-type arithmetic
 * (integer wrap-around; float sums carry the inter-tile prefix in f64).
 *
 * HBM traffic: numel * (sizeof(elem) + sizeof(sum)) bytes, read once, written once.
 */
#include "clo_internal.h"
#include "device_utils.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

using namespace clo;

namespace {

/* ------------------------------------------------------------------ traits */

template <typename SumT> struct AccOf {
	typedef typename std::conditional<std::is_floating_point<SumT>::value, double,
		typename std::conditional<sizeof(SumT) == 8, u64, u32>::type>::type type;
};

template <typename AccT> struct AccWords;
template <> struct AccWords<u32> {
	static const int N = 1;
	__device__ static void pack(u32 v, u32 (&w)[1]) { w[0] = v; }
	__device__ static u32 unpack(const u32 (&w)[1]) { return w[0]; }
};
template <> struct AccWords<u64> {
	static const int N = 2;
	__device__ static void pack(u64 v, u32 (&w)[2]) { w[0] = (u32) v; w[1] = (u32) (v >> 32); }
	__device__ static u64 unpack(const u32 (&w)[2]) { return (u64) w[0] | ((u64) w[1] << 32); }
};
template <> struct AccWords<double> {
	static const int N = 2;
	__device__ static void pack(double v, u32 (&w)[2]) {
		u64 b = (u64) __double_as_longlong(v); w[0] = (u32) b; w[1] = (u32) (b >> 32);
	}
	__device__ static double unpack(const u32 (&w)[2]) {
		return __longlong_as_double((long long) ((u64) w[0] | ((u64) w[1] << 32)));
	}
};

/* (SUM) in[j], widened to the accumulator: clo_scan_blelloch.cl:79-80 */
template <typename ElemT, typename SumT, typename AccT>
__device__ __forceinline__ AccT to_acc(ElemT e) {
	return static_cast<AccT>(static_cast<SumT>(e));
}

enum { ST_AGG = 1, ST_PREFIX = 2 };
const unsigned SPIN_LIMIT = 1u << 26;

/* ------------------------------------------------------------------ kernel */

/* One tile per CTA; tile index = blockIdx.x.  The look-back assumes that a CTA with a
 * lower index is never scheduled after one with a higher index (the order the hardware
 * work distributor uses for a 1-D grid, and the assumption behind cub::DeviceScan as
 * well); the spin bound turns a violation into an error flag instead of a hang.  Many
 * small CTAs per SM let the scheduler overlap one tile's look-back wait with the loads
 * of the others. */
template <typename ElemT, typename SumT, int THREADS, int VPT, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
clo_scan_lookback(const ElemT* __restrict__ in, SumT* __restrict__ out, size_t n,
		u64* __restrict__ desc, u32 epoch,
		const SumT* __restrict__ carry_in, int vec_in, int vec_out, int* __restrict__ err_flag) {
	typedef typename AccOf<SumT>::type AccT;
	typedef AccWords<AccT> AW;
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;       /* elements per vector */
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * VPT * EPV;
	/* output chunk: at most 16 bytes per store */
	constexpr int OCH = (sizeof(SumT) * EPV <= 16) ? EPV : (16 / (int) sizeof(SumT));

	__shared__ AccT s_warp[WARPS];
	__shared__ AccT s_prefix;

	const u32 tile = blockIdx.x;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const size_t tile_base = (size_t) tile * TILE;
	const bool full = tile_base + TILE <= n;

	/* ---- load: warp-striped vectors (row j of warp w is 32*EPV contiguous elements) */
	AccT v[VPT][EPV];
#pragma unroll
	for (int j = 0; j < VPT; ++j) {
		const size_t idx = tile_base + ((size_t) (warp * VPT + j) * 32 + lane) * EPV;
		ElemT e[EPV];
		if (vec_in && (full || idx + EPV <= n)) {
			load_vec_cs<ElemT, EPV>(in + idx, e);
		} else {
#pragma unroll
			for (int c = 0; c < EPV; ++c) e[c] = (idx + c < n) ? in[idx + c] : ElemT(0);
		}
#pragma unroll
		for (int c = 0; c < EPV; ++c) v[j][c] = to_acc<ElemT, SumT, AccT>(e[c]);
	}

	/* ---- thread: inclusive scan inside each vector; warp: scan of vector sums per row */
	AccT base[VPT];
	AccT warp_total = AccT(0);
#pragma unroll
	for (int j = 0; j < VPT; ++j) {
#pragma unroll
		for (int c = 1; c < EPV; ++c) v[j][c] += v[j][c - 1];
		AccT incl = warp_inclusive_scan<AccT>(v[j][EPV - 1], lane);
		AccT excl = __shfl_up_sync(0xffffffffu, incl, 1);
		if (lane == 0) excl = AccT(0);
		AccT row_total = __shfl_sync(0xffffffffu, incl, 31);
		base[j] = warp_total + excl;
		warp_total += row_total;
	}

	/* ---- block: offsets of the warps, tile aggregate */
	if (lane == 0) s_warp[warp] = warp_total;
	__syncthreads();
	AccT warp_off = AccT(0), aggregate = AccT(0);
#pragma unroll
	for (int w = 0; w < WARPS; ++w) {
		AccT t = s_warp[w];
		if (w < warp) warp_off += t;
		aggregate += t;
	}

	/* ---- decoupled look-back (warp 0; lane i inspects tile-1-i) */
	if (warp == 0) {
		const AccT carry = carry_in ? to_acc<SumT, SumT, AccT>(*carry_in) : AccT(0);
		const u64 fl_agg = ((u64) ((epoch << 2) | ST_AGG)) << 32;
		const u64 fl_pre = ((u64) ((epoch << 2) | ST_PREFIX)) << 32;
		u64* mine = desc + (size_t) tile * AW::N;
		AccT exclusive = carry;
		if (tile == 0) {
			if (lane == 0) {
				u32 w[AW::N];
				AW::pack(carry + aggregate, w);
#pragma unroll
				for (int k = 0; k < AW::N; ++k) st_relaxed(mine + k, fl_pre | w[k]);
			}
		} else {
			if (lane == 0) {
				u32 w[AW::N];
				AW::pack(aggregate, w);
#pragma unroll
				for (int k = 0; k < AW::N; ++k) st_relaxed(mine + k, fl_agg | w[k]);
			}
			exclusive = AccT(0);
			long long look = (long long) tile - 1;
			unsigned spins = 0;
			bool done = false;
			while (!done) {
				const long long idx = look - lane;
				u32 state = ST_PREFIX;  /* lanes before tile 0 act as an empty prefix */
				AccT val = AccT(0);
				if (idx >= 0) {
					const u64* p = desc + (size_t) idx * AW::N;
					for (;;) {
						u32 w[AW::N];
						u32 st = 0;
						bool ok = true;
#pragma unroll
						for (int k = 0; k < AW::N; ++k) {
							const u64 x = ld_relaxed(p + k);
							const u32 f = (u32) (x >> 32);
							w[k] = (u32) x;
							if ((f >> 2) != epoch || (f & 3u) == 0) ok = false;
							if (k == 0) st = f & 3u; else if ((f & 3u) != st) ok = false;
						}
						if (ok) { state = st; val = AW::unpack(w); break; }
						if (++spins > SPIN_LIMIT) { atomicExch(err_flag, 1); break; }
					}
				} else if (idx == -1) {
					val = carry;   /* the scan's carry-in sits "before tile 0" */
				}
				const u32 pmask = __ballot_sync(0xffffffffu, state == ST_PREFIX);
				const int first = pmask ? (__ffs(pmask) - 1) : 32;
				AccT contrib = (lane <= first) ? val : AccT(0);
				exclusive += warp_reduce_sum<AccT>(contrib);
				done = (pmask != 0);
				look -= 32;
			}
			if (lane == 0) {
				u32 w[AW::N];
				AW::pack(exclusive + aggregate, w);
#pragma unroll
				for (int k = 0; k < AW::N; ++k) st_relaxed(mine + k, fl_pre | w[k]);
			}
		}
		if (lane == 0) s_prefix = exclusive;
	}
	__syncthreads();
	const AccT tile_prefix = s_prefix + warp_off;

	/* ---- store: same warp-striped layout */
#pragma unroll
	for (int j = 0; j < VPT; ++j) {
		const size_t idx = tile_base + ((size_t) (warp * VPT + j) * 32 + lane) * EPV;
		const AccT b = tile_prefix + base[j];
		SumT o[EPV];
		o[0] = static_cast<SumT>(b);
#pragma unroll
		for (int c = 1; c < EPV; ++c) o[c] = static_cast<SumT>(b + v[j][c - 1]);
		if (vec_out && (full || idx + EPV <= n)) {
#pragma unroll
			for (int c0 = 0; c0 < EPV; c0 += OCH) {
				SumT chunk[OCH];
#pragma unroll
				for (int c = 0; c < OCH; ++c) chunk[c] = o[c0 + c];
				store_vec_cs<SumT, OCH>(out + idx + c0, chunk);
			}
		} else {
#pragma unroll
			for (int c = 0; c < EPV; ++c) if (idx + c < n) out[idx + c] = o[c];
		}
	}
}

/* Two-phase deterministic reduction (multi-GPU scan: per-GPU total). */
template <typename ElemT, typename SumT, int THREADS>
__global__ void __launch_bounds__(THREADS)
clo_scan_reduce_partial(const ElemT* __restrict__ in, size_t n, typename AccOf<SumT>::type* __restrict__ partial) {
	typedef typename AccOf<SumT>::type AccT;
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;
	__shared__ AccT s_warp[THREADS / 32];
	AccT acc = AccT(0);
	const size_t nvec = n / EPV;
	const bool vec_ok = (reinterpret_cast<uintptr_t>(in) % (sizeof(ElemT) * EPV)) == 0;
	if (vec_ok) {
		/* four independent 16-byte loads per thread and step; a step's 4 * EPV elements are summed in
		 * the tile-local type (f32 for f32 sums, like the scan kernels) before they join the f64 total */
		typedef typename std::conditional<std::is_same<SumT, float>::value, float, AccT>::type IntraT;
		const size_t stride = (size_t) gridDim.x * THREADS;
		size_t i = (size_t) blockIdx.x * THREADS + threadIdx.x;
		for (; i + 3 * stride < nvec; i += 4 * stride) {
			ElemT e[4][EPV];
#pragma unroll
			for (int u = 0; u < 4; ++u) load_vec_cs<ElemT, EPV>(in + (i + u * stride) * EPV, e[u]);
			IntraT part = IntraT(0);
#pragma unroll
			for (int u = 0; u < 4; ++u)
#pragma unroll
				for (int c = 0; c < EPV; ++c) part += to_acc<ElemT, SumT, IntraT>(e[u][c]);
			acc += static_cast<AccT>(part);
		}
		for (; i < nvec; i += stride) {
			ElemT e[EPV];
			load_vec_cs<ElemT, EPV>(in + i * EPV, e);
#pragma unroll
			for (int c = 0; c < EPV; ++c) acc += to_acc<ElemT, SumT, AccT>(e[c]);
		}
		for (size_t i = nvec * EPV + (size_t) blockIdx.x * THREADS + threadIdx.x; i < n; i += (size_t) gridDim.x * THREADS)
			acc += to_acc<ElemT, SumT, AccT>(in[i]);
	} else {
		for (size_t i = (size_t) blockIdx.x * THREADS + threadIdx.x; i < n; i += (size_t) gridDim.x * THREADS)
			acc += to_acc<ElemT, SumT, AccT>(in[i]);
	}
	acc = warp_reduce_sum<AccT>(acc);
	if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		AccT t = AccT(0);
		for (int w = 0; w < THREADS / 32; ++w) t += s_warp[w];
		partial[blockIdx.x] = t;
	}
}

template <typename SumT>
__global__ void clo_scan_reduce_final(const typename AccOf<SumT>::type* __restrict__ partial, int count,
		SumT* __restrict__ total_out) {
	typedef typename AccOf<SumT>::type AccT;
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		AccT t = AccT(0);
		for (int i = 0; i < count; ++i) t += partial[i];
		*total_out = static_cast<SumT>(t);
	}
}

#include "scan_pp.cuh"

/* --------------------------------------------------------------- host side */

struct ScanState {
	CloScratch scratch;        /* [pad | err(int) | pad | descriptors...] */
	size_t tiles_cap = 0;
	u32 epoch = 0;
	int cfg = 0;               /* CLO_SCAN_CFG: tile shape variants of the u32->u32 kernel */
	CloScratch partials;
	CloScratch pp;             /* persistent kernel: [ticket | pad][AGG words][PREF words] */
	size_t pp_tiles_cap = 0;
	u32 pp_epoch = 0;
	int use_pp = 1;            /* CLO_SCAN_KERNEL=classic selects the one-tile-per-CTA kernel */
	int tma_flags = 0;         /* CLO_SCAN_TMA_FLAGS: 1 windowed propagator, 2 chain with 64-tile chunks, 4 no poll back-off */
	int use_tma = 1;           /* CLO_SCAN_KERNEL=tma: the copy-engine kernel for same-size types (scan_tma.cuh) */
};

const size_t HDR_BYTES = 256;

/* default tile: 128 threads x 4 vectors (2048 4-byte elements), >= 10 CTAs per SM */
const int SCAN_THREADS = 128;
const int SCAN_VPT = 4;
const int SCAN_MIN_CTAS = 10;

template <typename ElemT, typename SumT, int THREADS, int VPT, int MIN_CTAS>
cudaError_t launch_scan_cfg(ScanState& st, const void* in, void* out, size_t n, const void* carry, cudaStream_t stream) {
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;
	constexpr size_t TILE = (size_t) THREADS * VPT * EPV;
	const size_t tiles = (n + TILE - 1) / TILE;
	if (tiles >= 0x7fffffffull) return cudaErrorInvalidValue;
	cudaError_t e;
	if (tiles > st.tiles_cap) {
		/* descriptors are sized for the widest accumulator (2 words) */
		size_t cap = tiles + tiles / 4 + 1024;
		if ((e = st.scratch.reserve(HDR_BYTES + cap * 2 * sizeof(u64))) != cudaSuccess) return e;
		if ((e = cudaMemsetAsync(st.scratch.ptr, 0, st.scratch.size, stream)) != cudaSuccess) return e;
		st.tiles_cap = cap; st.epoch = 0;
	}
	if (st.epoch >= (1u << 30) - 2) {
		if ((e = cudaMemsetAsync(st.scratch.ptr, 0, st.scratch.size, stream)) != cudaSuccess) return e;
		st.epoch = 0;
	}
	/* descriptors of earlier calls carry an older epoch and read as "not published" */
	st.epoch += 1;
	int* err_flag = (int*) st.scratch.ptr + 1;
	u64* desc = (u64*) ((char*) st.scratch.ptr + HDR_BYTES);
	const int vec_in = (reinterpret_cast<uintptr_t>(in) % (sizeof(ElemT) * EPV)) == 0;
	constexpr int OCH = (sizeof(SumT) * EPV <= 16) ? EPV : (16 / (int) sizeof(SumT));
	const int vec_out = (reinterpret_cast<uintptr_t>(out) % (sizeof(SumT) * OCH)) == 0;
	clo_scan_lookback<ElemT, SumT, THREADS, VPT, MIN_CTAS><<<(unsigned) tiles, THREADS, 0, stream>>>(
		(const ElemT*) in, (SumT*) out, n, desc, st.epoch, (const SumT*) carry, vec_in, vec_out, err_flag);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

template <typename ElemT, typename SumT>
cudaError_t launch_scan(ScanState& st, const void* in, void* out, size_t n, const void* carry, int sms, cudaStream_t stream) {
	(void) sms;
	return launch_scan_cfg<ElemT, SumT, SCAN_THREADS, SCAN_VPT, SCAN_MIN_CTAS>(st, in, out, n, carry, stream);
}

/* persistent kernel (scan_pp.cuh): used for large, 16-byte aligned inputs of the hot type pairs */
const size_t SPP_MIN_ELEMS = (size_t) 1 << 22;

template <typename ElemT, typename SumT, int THREADS = SPP_THREADS, int VPT = SPP_VPT, int AHEAD = SPP_AHEAD, int LAG = SPP_LAG, bool ONEBAR = false>
cudaError_t launch_scan_pp(ScanState& st, const void* in, void* out, size_t n, const void* carry, int sms, cudaStream_t stream) {
	typedef typename AccOf<SumT>::type AccT;
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;
	constexpr size_t TILE = (size_t) THREADS * VPT * EPV;
	constexpr size_t SMEM = (size_t) (AHEAD + 1 + LAG) * TILE * sizeof(ElemT);
	const size_t tiles = (n + TILE - 1) / TILE;
	cudaError_t e;
	if (tiles > st.pp_tiles_cap) {
		size_t cap = tiles + tiles / 4 + 1024;
		if ((e = st.pp.reserve(HDR_BYTES + cap * 4 * sizeof(u64))) != cudaSuccess) return e;
		if ((e = cudaMemsetAsync(st.pp.ptr, 0, st.pp.size, stream)) != cudaSuccess) return e;
		st.pp_tiles_cap = cap; st.pp_epoch = 0;
	}
	if (st.pp_epoch >= 0xfffffff0u) {
		if ((e = cudaMemsetAsync(st.pp.ptr, 0, st.pp.size, stream)) != cudaSuccess) return e;
		st.pp_epoch = 0;
	}
	st.pp_epoch += 1;
	u32* ticket = (u32*) st.pp.ptr;
	int* err_flag = (int*) st.pp.ptr + 1;
	u64* agg = (u64*) ((char*) st.pp.ptr + HDR_BYTES);
	u64* pref = agg + st.pp_tiles_cap * 2;
	if ((e = cudaMemsetAsync(ticket, 0, sizeof(u32), stream)) != cudaSuccess) return e;
	auto kern = ONEBAR ? clo_scan_pp1b<ElemT, SumT, THREADS, VPT, AHEAD, LAG> : clo_scan_pp<ElemT, SumT, THREADS, VPT, AHEAD, LAG>;
	/* the shared-memory opt-in and the occupancy are per-device state */
	static bool configured[64] = {};
	static int ctas_per_sm_dev[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	if (!configured[dev]) {
		if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM)) != cudaSuccess) return e;
		int k = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, THREADS, SMEM) != cudaSuccess || k < 1) k = 1;
		ctas_per_sm_dev[dev] = k;
		configured[dev] = true;
	}
	const int ctas_per_sm = ctas_per_sm_dev[dev];
	size_t workers = (size_t) sms * ctas_per_sm - 1;
	if (workers > tiles) workers = tiles;
	kern<<<(unsigned) (1 + workers), THREADS, SMEM, stream>>>((const ElemT*) in, (SumT*) out, n, (u32) tiles,
		agg, pref, ticket, st.pp_epoch, (const SumT*) carry, err_flag);
	CLO_COUNT_LAUNCH(1);
	(void) sizeof(AccT);
	return cudaGetLastError();
}

#include "scan_tma.cuh"

typedef CUresult (*StmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
	const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static StmEncodeFn stm_encode_fn() {
	static StmEncodeFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void* p = nullptr;
		cudaDriverEntryPointQueryResult qr;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
				qr == cudaDriverEntryPointSuccess) fn = (StmEncodeFn) p;
	}
	return fn;
}

/* a buffer of `bytes` bytes (a multiple of 128) as [bytes / 128] rows of 32 words; box = one tile */
static bool stm_make_map(CUtensorMap* tm, const void* ptr, size_t bytes, int box_rows) {
	StmEncodeFn enc = stm_encode_fn();
	if (!enc) return false;
	const cuuint64_t dims[2] = { 32, (cuuint64_t) (bytes / 128) };
	const cuuint64_t strides[1] = { 128 };
	const cuuint32_t box[2] = { 32, (cuuint32_t) box_rows };
	const cuuint32_t estr[2] = { 1, 1 };
	return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
		CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename ElemT, typename SumT>
bool scan_tma_applicable(const ScanState& st, const void* in, const void* out, size_t n) {
	return st.use_pp && st.use_tma && sizeof(ElemT) == sizeof(SumT) && n >= SPP_MIN_ELEMS &&
		(n * sizeof(ElemT)) % 128 == 0 && (n * sizeof(ElemT)) / 128 < 0x7fffffffull &&
		(reinterpret_cast<uintptr_t>(in) % 16) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0 && stm_encode_fn() != nullptr;
}

template <typename ElemT, typename SumT, int THREADS = 256, int AHEAD = 1, int LAG = 4, int SLACK = 0>
cudaError_t launch_scan_tma(ScanState& st, const void* in, void* out, size_t n, const void* carry, int sms, cudaStream_t stream) {
	typedef typename AccOf<SumT>::type AccT;
	typedef typename std::conditional<std::is_same<SumT, float>::value, float, AccT>::type IntraT;
	typedef StmShape<THREADS, AHEAD, LAG, SLACK> Shape;
	constexpr size_t TILE = (size_t) Shape::TILE_BYTES / sizeof(ElemT);
	constexpr size_t SMEM = 1024 + (size_t) Shape::S * Shape::TILE_BYTES + (size_t) Shape::S * THREADS * sizeof(IntraT);
	const size_t tiles = (n + TILE - 1) / TILE;
	cudaError_t e;
	CUtensorMap tm_in, tm_out;
	if (!stm_make_map(&tm_in, in, n * sizeof(ElemT), Shape::ROWS) || !stm_make_map(&tm_out, out, n * sizeof(SumT), Shape::ROWS))
		return cudaErrorInvalidValue;
	if (tiles > st.pp_tiles_cap) {
		size_t cap = tiles + tiles / 4 + 1024;
		if ((e = st.pp.reserve(HDR_BYTES + cap * 4 * sizeof(u64))) != cudaSuccess) return e;
		if ((e = cudaMemsetAsync(st.pp.ptr, 0, st.pp.size, stream)) != cudaSuccess) return e;
		st.pp_tiles_cap = cap; st.pp_epoch = 0;
	}
	if (st.pp_epoch >= 0xfffffff0u) {
		if ((e = cudaMemsetAsync(st.pp.ptr, 0, st.pp.size, stream)) != cudaSuccess) return e;
		st.pp_epoch = 0;
	}
	st.pp_epoch += 1;
	u32* ticket = (u32*) st.pp.ptr;
	int* err_flag = (int*) st.pp.ptr + 1;
	u64* agg = (u64*) ((char*) st.pp.ptr + HDR_BYTES);
	u64* pref = agg + st.pp_tiles_cap * 2;
	if ((e = cudaMemsetAsync(ticket, 0, sizeof(u32), stream)) != cudaSuccess) return e;
	auto kern = clo_scan_tma<ElemT, SumT, THREADS, AHEAD, LAG, SLACK>;
	static bool configured[64] = {};
	static int ctas_per_sm_dev[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64) dev = 0;
	if (!configured[dev]) {
		if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM)) != cudaSuccess) return e;
		int k = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&k, kern, THREADS, SMEM) != cudaSuccess || k < 1) k = 1;
		ctas_per_sm_dev[dev] = k;
		configured[dev] = true;
	}
	size_t workers = (size_t) sms * ctas_per_sm_dev[dev] - 1;
	if (workers > tiles) workers = tiles;
	kern<<<(unsigned) (1 + workers), THREADS, SMEM, stream>>>(tm_in, tm_out, (u32) tiles,
		agg, pref, ticket, st.pp_epoch, (const SumT*) carry, err_flag, st.tma_flags);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

template <typename ElemT, typename SumT>
bool scan_pp_applicable(const ScanState& st, const void* in, const void* out, size_t n) {
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;
	constexpr int OCH = (sizeof(SumT) * EPV <= 16) ? EPV : (16 / (int) sizeof(SumT));
	return st.use_pp && n >= SPP_MIN_ELEMS && (reinterpret_cast<uintptr_t>(in) % 16) == 0 &&
		(reinterpret_cast<uintptr_t>(out) % (sizeof(SumT) * OCH)) == 0;
}

#define CLO_SCAN_PP_PAIR(E, S) \
template <> \
cudaError_t launch_scan<E, S>(ScanState& st, const void* in, void* out, size_t n, const void* carry, int sms, cudaStream_t stream) { \
	if (scan_pp_applicable<E, S>(st, in, out, n)) return launch_scan_pp<E, S>(st, in, out, n, carry, sms, stream); \
	return launch_scan_cfg<E, S, SCAN_THREADS, SCAN_VPT, SCAN_MIN_CTAS>(st, in, out, n, carry, stream); \
}
/* same-size pairs: the copy-engine kernel first (scan_tma.cuh), then the cp.async ring */
#define CLO_SCAN_TMA_PAIR(E, S) \
template <> \
cudaError_t launch_scan<E, S>(ScanState& st, const void* in, void* out, size_t n, const void* carry, int sms, cudaStream_t stream) { \
	if (scan_tma_applicable<E, S>(st, in, out, n)) return launch_scan_tma<E, S>(st, in, out, n, carry, sms, stream); \
	if (scan_pp_applicable<E, S>(st, in, out, n)) return launch_scan_pp<E, S>(st, in, out, n, carry, sms, stream); \
	return launch_scan_cfg<E, S, SCAN_THREADS, SCAN_VPT, SCAN_MIN_CTAS>(st, in, out, n, carry, stream); \
}
CLO_SCAN_PP_PAIR(unsigned int, unsigned long long)
CLO_SCAN_TMA_PAIR(int, int)
CLO_SCAN_PP_PAIR(int, long long)

CLO_SCAN_PP_PAIR(float, double)
CLO_SCAN_TMA_PAIR(unsigned long long, unsigned long long)
CLO_SCAN_TMA_PAIR(long long, long long)
CLO_SCAN_TMA_PAIR(double, double)
#undef CLO_SCAN_PP_PAIR
#undef CLO_SCAN_TMA_PAIR

/* tuning variants of the headline type pair */
template <>
cudaError_t launch_scan<unsigned int, unsigned int>(ScanState& st, const void* in, void* out, size_t n,
		const void* carry, int sms, cudaStream_t stream) {
	if (scan_tma_applicable<unsigned int, unsigned int>(st, in, out, n)) {
		if (st.cfg == 32) return launch_scan_tma<unsigned int, unsigned int, 256, 1, 3, 1>(st, in, out, n, carry, sms, stream);
		return launch_scan_tma<unsigned int, unsigned int>(st, in, out, n, carry, sms, stream);
	}
	if (scan_pp_applicable<unsigned int, unsigned int>(st, in, out, n)) {
		switch (st.cfg) {      /* CLO_SCAN_CFG 10..: shapes of the persistent kernel */
		case 11: return launch_scan_pp<unsigned int, unsigned int, 512, 2, 1, 4>(st, in, out, n, carry, sms, stream);
		case 12: return launch_scan_pp<unsigned int, unsigned int, 256, 2, 1, 8>(st, in, out, n, carry, sms, stream);
		case 13: return launch_scan_pp<unsigned int, unsigned int, 128, 4, 1, 6>(st, in, out, n, carry, sms, stream);
		case 14: return launch_scan_pp<unsigned int, unsigned int, 256, 3, 1, 6>(st, in, out, n, carry, sms, stream);
		case 15: return launch_scan_pp<unsigned int, unsigned int, 512, 4, 1, 1>(st, in, out, n, carry, sms, stream);
		case 20: return launch_scan_pp<unsigned int, unsigned int, 256, 4, 1, 4, true>(st, in, out, n, carry, sms, stream);
		case 21: return launch_scan_pp<unsigned int, unsigned int, 256, 3, 1, 6, true>(st, in, out, n, carry, sms, stream);
		case 22: return launch_scan_pp<unsigned int, unsigned int, 256, 4, 1, 5, true>(st, in, out, n, carry, sms, stream);
		case 23: return launch_scan_pp<unsigned int, unsigned int, 128, 4, 1, 6, true>(st, in, out, n, carry, sms, stream);
		default: return launch_scan_pp<unsigned int, unsigned int>(st, in, out, n, carry, sms, stream);
		}
	}
	switch (st.cfg) {
	case 1: return launch_scan_cfg<unsigned int, unsigned int, 256, 4, 5>(st, in, out, n, carry, stream);
	case 2: return launch_scan_cfg<unsigned int, unsigned int, 128, 8, 6>(st, in, out, n, carry, stream);
	case 3: return launch_scan_cfg<unsigned int, unsigned int, 64, 4, 16>(st, in, out, n, carry, stream);
	case 4: return launch_scan_cfg<unsigned int, unsigned int, 256, 8, 3>(st, in, out, n, carry, stream);
	case 5: return launch_scan_cfg<unsigned int, unsigned int, 128, 4, 16>(st, in, out, n, carry, stream);
	default: return launch_scan_cfg<unsigned int, unsigned int, SCAN_THREADS, SCAN_VPT, SCAN_MIN_CTAS>(st, in, out, n, carry, stream);
	}
}

template <>
cudaError_t launch_scan<float, float>(ScanState& st, const void* in, void* out, size_t n,
		const void* carry, int sms, cudaStream_t stream) {
	if (scan_tma_applicable<float, float>(st, in, out, n)) {
		if (st.cfg == 32) return launch_scan_tma<float, float, 256, 1, 3, 1>(st, in, out, n, carry, sms, stream);
		return launch_scan_tma<float, float>(st, in, out, n, carry, sms, stream);
	}
	if (scan_pp_applicable<float, float>(st, in, out, n)) {
		switch (st.cfg) {      /* CLO_SCAN_CFG 10..: shapes of the persistent kernel */
		case 11: return launch_scan_pp<float, float, 512, 2, 1, 4>(st, in, out, n, carry, sms, stream);
		case 12: return launch_scan_pp<float, float, 256, 2, 1, 8>(st, in, out, n, carry, sms, stream);
		case 13: return launch_scan_pp<float, float, 128, 4, 1, 6>(st, in, out, n, carry, sms, stream);
		case 14: return launch_scan_pp<float, float, 256, 3, 1, 6>(st, in, out, n, carry, sms, stream);
		case 15: return launch_scan_pp<float, float, 512, 2, 2, 3>(st, in, out, n, carry, sms, stream);
		case 16: return launch_scan_pp<float, float, 1024, 1, 1, 4>(st, in, out, n, carry, sms, stream);
		case 20: return launch_scan_pp<float, float, 256, 4, 1, 4, true>(st, in, out, n, carry, sms, stream);
		case 21: return launch_scan_pp<float, float, 256, 3, 1, 6, true>(st, in, out, n, carry, sms, stream);
		case 22: return launch_scan_pp<float, float, 256, 4, 1, 5, true>(st, in, out, n, carry, sms, stream);
		case 23: return launch_scan_pp<float, float, 128, 4, 1, 6, true>(st, in, out, n, carry, sms, stream);
		case 24: return launch_scan_pp<float, float, 128, 4, 1, 4, true>(st, in, out, n, carry, sms, stream);
		case 25: return launch_scan_pp<float, float, 256, 2, 1, 8, true>(st, in, out, n, carry, sms, stream);
		default: return launch_scan_pp<float, float>(st, in, out, n, carry, sms, stream);
		}
	}
	return launch_scan_cfg<float, float, SCAN_THREADS, SCAN_VPT, SCAN_MIN_CTAS>(st, in, out, n, carry, stream);
}

template <typename ElemT, typename SumT>
cudaError_t launch_reduce(ScanState& st, const void* in, void* total_out, size_t n, int sms, cudaStream_t stream) {
	typedef typename AccOf<SumT>::type AccT;
	const int blocks = sms * 8;
	cudaError_t e;
	if ((e = st.partials.reserve((size_t) blocks * sizeof(AccT))) != cudaSuccess) return e;
	clo_scan_reduce_partial<ElemT, SumT, 256><<<blocks, 256, 0, stream>>>((const ElemT*) in, n, (AccT*) st.partials.ptr);
	clo_scan_reduce_final<SumT><<<1, 32, 0, stream>>>((const AccT*) st.partials.ptr, blocks, (SumT*) total_out);
	CLO_COUNT_LAUNCH(2);
	return cudaGetLastError();
}

/* carry of the NEXT chunk of a chunked scan: exclusive sum of the chunk's last element + that element */
template <typename ElemT, typename SumT>
__global__ void clo_scan_next_carry(const ElemT* __restrict__ in, const SumT* __restrict__ out, size_t n, SumT* __restrict__ carry) {
	if (threadIdx.x == 0 && blockIdx.x == 0 && n) *carry = static_cast<SumT>(out[n - 1] + static_cast<SumT>(in[n - 1]));
}
template <typename ElemT, typename SumT>
cudaError_t launch_next_carry(const void* in, const void* out, size_t n, void* carry, cudaStream_t stream) {
	clo_scan_next_carry<ElemT, SumT><<<1, 32, 0, stream>>>((const ElemT*) in, (const SumT*) out, n, (SumT*) carry);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

typedef cudaError_t (*ScanFn)(ScanState&, const void*, void*, size_t, const void*, int, cudaStream_t);
typedef cudaError_t (*ReduceFn)(ScanState&, const void*, void*, size_t, int, cudaStream_t);
typedef cudaError_t (*CarryFn)(const void*, const void*, size_t, void*, cudaStream_t);

/* CloType -> C++ type (half is not an arithmetic type in OpenCL C without
 * cl_khr_fp16; unsupported here as well) */
template <int T> struct CT;
template <> struct CT<CLO_CHAR> { typedef signed char type; };
template <> struct CT<CLO_UCHAR> { typedef unsigned char type; };
template <> struct CT<CLO_SHORT> { typedef short type; };
template <> struct CT<CLO_USHORT> { typedef unsigned short type; };
template <> struct CT<CLO_INT> { typedef int type; };
template <> struct CT<CLO_UINT> { typedef unsigned int type; };
template <> struct CT<CLO_LONG> { typedef long long type; };
template <> struct CT<CLO_ULONG> { typedef unsigned long long type; };
template <> struct CT<CLO_FLOAT> { typedef float type; };
template <> struct CT<CLO_DOUBLE> { typedef double type; };

template <int E, int S> struct Entry {
	static ScanFn scan() { return &launch_scan<typename CT<E>::type, typename CT<S>::type>; }
	static ReduceFn reduce() { return &launch_reduce<typename CT<E>::type, typename CT<S>::type>; }
	static CarryFn carry() { return &launch_next_carry<typename CT<E>::type, typename CT<S>::type>; }
};

#define CLO_TYPE_CASES(M) \
	M(CLO_CHAR) M(CLO_UCHAR) M(CLO_SHORT) M(CLO_USHORT) M(CLO_INT) M(CLO_UINT) \
	M(CLO_LONG) M(CLO_ULONG) M(CLO_FLOAT) M(CLO_DOUBLE)

template <int E> ScanFn scan_for_sum(int s) {
	switch (s) {
#define CLO_M(S) case S: return Entry<E, S>::scan();
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

template <int E> ReduceFn reduce_for_sum(int s) {
	switch (s) {
#define CLO_M(S) case S: return Entry<E, S>::reduce();
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

ScanFn find_scan(int e, int s) {
	switch (e) {
#define CLO_M(E) case E: return scan_for_sum<E>(s);
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

template <int E> CarryFn carry_for_sum(int s) {
	switch (s) {
#define CLO_M(S) case S: return Entry<E, S>::carry();
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

CarryFn find_carry(int e, int s) {
	switch (e) {
#define CLO_M(E) case E: return carry_for_sum<E>(s);
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

ReduceFn find_reduce(int e, int s) {
	switch (e) {
#define CLO_M(E) case E: return reduce_for_sum<E>(s);
	CLO_TYPE_CASES(CLO_M)
#undef CLO_M
	default: return nullptr;
	}
}

} // namespace

/* ------------------------------------------------------------------ object */

struct clo_scan {
	CloScanImplDef impl_def;
	CCLContext* ctx;
	CCLProgram* prg;
	CloType elem_type;
	CloType sum_type;
	void* data;
	/* backend state */
	ScanState st;
	ScanFn fn;
	ReduceFn rfn;
	CloScratch host_in, host_out;   /* device staging of clo_scan_with_host_data, kept between calls */
	CarryFn cfn;
	cudaStream_t s_in = nullptr, s_out = nullptr;      /* copy streams of the chunked host path */
	cudaEvent_t ev_in[4] = {}, ev_sc[4] = {};
	void* d_chunk_carry = nullptr;
};

static ccl_program g_scan_program = { "clo_scan (precompiled sm_100a)", nullptr, std::string(), nullptr, {} };

static const char* blelloch_init(CloScan* scanner, const char* options, GError** err) {
	(void) scanner;
	/* clo_scan_blelloch.c:43-45: no options are accepted */
	if (options && *options) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Invalid options for blelloch scan.");
		return NULL;
	}
	return "";
}

static void blelloch_finalize(CloScan* scanner) { (void) scanner; }

static int scan_check_error_flag(CloScan* scanner, cudaStream_t stream, GError** err) {
	if (!scanner->st.scratch.ptr && !scanner->st.pp.ptr) return 0;
	int flag = 0, flag2 = 0;
	if (scanner->st.scratch.ptr && clo_cuda_failed(cudaMemcpyAsync(&flag, (int*) scanner->st.scratch.ptr + 1, sizeof(int),
			cudaMemcpyDeviceToHost, stream), err, "scan status read")) return 1;
	if (scanner->st.pp.ptr && clo_cuda_failed(cudaMemcpyAsync(&flag2, (int*) scanner->st.pp.ptr + 1, sizeof(int),
			cudaMemcpyDeviceToHost, stream), err, "scan status read")) return 1;
	if (clo_cuda_failed(cudaStreamSynchronize(stream), err, "scan sync")) return 1;
	flag |= flag2;
	if (flag) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "scan look-back timed out (device status flag set)");
		return 1;
	}
	return 0;
}

static CCLEvent* scan_device(CloScan* scanner, CCLQueue* cq_exec, CCLBuffer* data_in,
		CCLBuffer* data_out, CCLBuffer* carry, size_t numel, GError** err) {
	if (!scanner || !cq_exec || !data_in || !data_out) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "scan: NULL argument");
		return NULL;
	}
	const size_t es = clo_type_sizeof(scanner->elem_type), ss = clo_type_sizeof(scanner->sum_type);
	if (data_in->size < numel * es || data_out->size < numel * ss) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "scan: buffers too small for %zu elements", numel);
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_scan_lookback");
	cudaError_t rc = cudaSuccess;
	if (numel > 0)
		rc = scanner->fn(scanner->st, data_in->ptr, data_out->ptr, numel, carry ? carry->ptr : NULL,
			clo_sm_count(cq_exec->ctx->dev.ordinal), cq_exec->stream);
	clo_queue_end(cq_exec, evt);
	if (clo_cuda_failed(rc, err, "clo_scan_lookback launch")) return NULL;
	return evt;
}

static CCLEvent* blelloch_scan_with_device_data(CloScan* scanner, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	(void) cq_comm; (void) lws_max;
	return scan_device(scanner, cq_exec, data_in, data_out, NULL, numel, err);
}

static const char* const kScanKernels[] = { "clo_scan_tma", "clo_scan_pp", "clo_scan_lookback", "clo_scan_reduce_partial", "clo_scan_reduce_final" };

static cl_uint blelloch_get_num_kernels(CloScan* scanner, GError** err) { (void) scanner; (void) err; return 5; }

static const char* blelloch_get_kernel_name(CloScan* scanner, cl_uint i, GError** err) {
	(void) scanner;
	if (i >= 5) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return NULL; }
	return kScanKernels[i];
}

static size_t blelloch_get_localmem_usage(CloScan* scanner, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	(void) lws_max; (void) numel;
	if (i >= 5) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return 0; }
	const size_t acc = clo_type_sizeof(scanner->sum_type) == 8 || scanner->sum_type >= CLO_FLOAT ? 8 : 4;
	if (i == 0) {
		/* clo_scan_tma: ring of AHEAD + LAG + 1 slots of 256 x 64 bytes (+ 1 KB alignment), the lanes-
		 * below sums per slot, mbarriers, tickets, warp totals */
		typedef StmShape<256, 1, 4, 0> Shape;
		const size_t intra = scanner->sum_type == CLO_FLOAT ? 4 : acc;
		return 1024 + (size_t) Shape::S * Shape::TILE_BYTES + (size_t) Shape::S * 256 * intra +
			Shape::S * 8 + (Shape::S + 2) * 4 + (2 + Shape::S) * 8 * acc;
	}
	i -= 1;
	if (i == 0) {
		/* clo_scan_pp: ring of AHEAD + 1 + LAG tiles of 256 threads x 4 vectors of 16 bytes, the
		 * per-slot warp totals and the ticket ring */
		const size_t slots = SPP_AHEAD + 1 + SPP_LAG;
		return slots * (size_t) SPP_THREADS * SPP_VPT * 16 + slots * (SPP_THREADS / 32) * acc + (slots + 2) * 4;
	}
	if (i == 1) return 8 + (SCAN_THREADS / 32 + 1) * acc;       /* clo_scan_lookback: tile id + warp totals + tile prefix */
	return i == 2 ? 8 * acc : 0;                                   /* clo_scan_reduce_partial: warp totals */
}

extern "C" const CloScanImplDef clo_scan_blelloch_def = {
	"blelloch", blelloch_init, blelloch_finalize, blelloch_scan_with_device_data,
	blelloch_get_num_kernels, blelloch_get_kernel_name, blelloch_get_localmem_usage
};

extern "C" CloScan* clo_scan_new(const char* type, const char* options, CCLContext* ctx,
		CloType elem_type, CloType sum_type, const char* compiler_opts, GError** err) {
	(void) compiler_opts; /* OpenCL build options have no meaning here; accepted and ignored */
	if (err && *err) return NULL;
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	if (!type || strcmp(type, clo_scan_blelloch_def.name) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_IMPL_NOT_FOUND,
			"The requested scan implementation, '%s', was not found.", type ? type : "(null)");
		return NULL;
	}
	ScanFn fn = find_scan((int) elem_type, (int) sum_type);
	if (!fn) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_UNKNOWN_TYPE,
			"Unsupported scan types (elem=%d, sum=%d)", (int) elem_type, (int) sum_type);
		return NULL;
	}
	clo_scan* s = new clo_scan();
	s->impl_def = clo_scan_blelloch_def;
	s->ctx = ctx; ccl_context_ref(ctx);
	s->prg = &g_scan_program;
	s->elem_type = elem_type; s->sum_type = sum_type;
	s->data = NULL;
	s->fn = fn; s->rfn = find_reduce((int) elem_type, (int) sum_type); s->cfn = find_carry((int) elem_type, (int) sum_type);
	{ const char* c = getenv("CLO_SCAN_CFG"); s->st.cfg = (c && *c) ? atoi(c) : 0; }
	{ const char* c = getenv("CLO_SCAN_KERNEL"); s->st.use_pp = (c && strcmp(c, "classic") == 0) ? 0 : 1;
	  s->st.use_tma = (c && (strcmp(c, "pp") == 0 || strcmp(c, "classic") == 0)) ? 0 : 1;
	  const char* f = getenv("CLO_SCAN_TMA_FLAGS"); s->st.tma_flags = (f && *f) ? atoi(f) : 0; }
	GError* ierr = NULL;
	s->impl_def.init(s, options, &ierr);
	if (ierr) { g_propagate_error(err, ierr); clo_scan_destroy(s); return NULL; }
	clo_handle_add(s);
	return s;
}

extern "C" void clo_scan_destroy(CloScan* scan) {
	if (!scan) return;
	clo_handle_remove(scan);
	scan->impl_def.finalize(scan);
	{
		CloDeviceGuard g(scan->ctx->dev.ordinal);
		if (getenv("CLO_SCAN_STATS") && scan->st.pp.ptr) {
			/* development aid: prefix-word polls of the persistent kernel since the scanner was made */
			unsigned h[4] = {};
			cudaDeviceSynchronize();
			cudaMemcpy(h, scan->st.pp.ptr, sizeof(h), cudaMemcpyDeviceToHost);
			fprintf(stderr, "clo_scan stats: err %u prefix polls %u\n", h[1], h[3]);
		}
		scan->st.scratch.release();
		scan->st.partials.release();
		scan->st.pp.release();
		scan->host_in.release();
		scan->host_out.release();
		if (scan->s_in) cudaStreamDestroy(scan->s_in);
		if (scan->s_out) cudaStreamDestroy(scan->s_out);
		for (int i = 0; i < 4; ++i) { if (scan->ev_in[i]) cudaEventDestroy(scan->ev_in[i]); if (scan->ev_sc[i]) cudaEventDestroy(scan->ev_sc[i]); }
		if (scan->d_chunk_carry) cudaFree(scan->d_chunk_carry);
	}
	ccl_context_unref(scan->ctx);
	delete scan;
}

extern "C" CCLEvent* clo_scan_with_device_data(CloScan* scanner, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	if (!scanner || (err && *err) || !cq_exec) return NULL;
	return scanner->impl_def.scan_with_device_data(scanner, cq_exec, cq_comm, data_in, data_out, numel, lws_max, err);
}

extern "C" CCLEvent* clo_scan_with_device_data_carry(CloScan* scanner, CCLQueue* cq_exec,
		CCLBuffer* data_in, CCLBuffer* data_out, CCLBuffer* carry_in, size_t numel, GError** err) {
	if (!scanner || (err && *err) || !cq_exec) return NULL;
	return scan_device(scanner, cq_exec, data_in, data_out, carry_in, numel, err);
}

extern "C" CCLEvent* clo_scan_reduce_with_device_data(CloScan* scanner, CCLQueue* cq_exec,
		CCLBuffer* data_in, CCLBuffer* total_out, size_t numel, GError** err) {
	if (!scanner || (err && *err) || !cq_exec || !data_in || !total_out) return NULL;
	if (data_in->size < numel * clo_type_sizeof(scanner->elem_type) ||
			total_out->size < clo_type_sizeof(scanner->sum_type)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "scan reduce: buffers too small");
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_scan_reduce");
	cudaError_t rc = scanner->rfn(scanner->st, data_in->ptr, total_out->ptr, numel,
		clo_sm_count(cq_exec->ctx->dev.ordinal), cq_exec->stream);
	clo_queue_end(cq_exec, evt);
	if (clo_cuda_failed(rc, err, "clo_scan_reduce launch")) return NULL;
	return evt;
}

/* clo_scan_abstract.c:255-362: alloc in/out, H2D, scan, D2H, block */
/* Host data in, scanned host data out (clo_scan_abstract.c:255-362: blocks until data_out is
 * complete).  Device staging is kept in the scanner between calls; with one queue the copy in, the
 * scan and the copy out are stream ordered and the host blocks once. */
extern "C" cl_bool clo_scan_with_host_data(CloScan* scanner, CCLQueue* cq_exec, CCLQueue* cq_comm,
		void* data_in, void* data_out, size_t numel, size_t lws_max, GError** err) {
	if (!scanner || (err && *err)) return CL_FALSE;
	if (numel && (!data_in || !data_out)) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "scan: NULL host pointer"); return CL_FALSE; }
	const size_t in_size = numel * clo_type_sizeof(scanner->elem_type);
	const size_t out_size = numel * clo_type_sizeof(scanner->sum_type);
	CCLQueue* own_queue = NULL;
	if (!cq_exec) {
		own_queue = ccl_queue_new(scanner->ctx, NULL, 0, err);
		if (!own_queue) return CL_FALSE;
		cq_exec = own_queue;
	}
	if (!cq_comm) cq_comm = cq_exec;
	const bool one_queue = cq_comm == cq_exec;
	cl_bool ok = CL_FALSE;
	CCLBuffer *in_dev = NULL, *out_dev = NULL;
	CCLEventWaitList ewl = NULL;
	{
		CloDeviceGuard g(scanner->ctx->dev.ordinal);
		if (clo_cuda_failed(scanner->host_in.reserve(in_size ? in_size : 1), err, "scan staging") ||
				clo_cuda_failed(scanner->host_out.reserve(out_size ? out_size : 1), err, "scan staging")) goto done;
		in_dev = ccl_buffer_new_wrap(scanner->ctx, scanner->host_in.ptr, in_size, err);
		if (in_dev) out_dev = ccl_buffer_new_wrap(scanner->ctx, scanner->host_out.ptr, out_size, err);
		if (!in_dev || !out_dev) goto done;
		/* Large inputs go through in CHUNKS: while chunk i is scanned (carry-in = everything before
		 * it, kept on the device), chunk i+1 is on its way in and chunk i-1 on its way out -- the two
		 * directions of the host link run at the same time, so the call costs about ONE transfer of
		 * the larger side instead of copy in + scan + copy out.  CLO_SCAN_HOST_CHUNK (elements) tunes
		 * or disables (0) it. */
		const char* chunk_e = getenv("CLO_SCAN_HOST_CHUNK");
		const size_t chunk = (chunk_e && *chunk_e) ? (size_t) atoll(chunk_e) : ((size_t) 1 << 24);
		if (one_queue && chunk && numel >= 2 * chunk && scanner->cfn) {
			const size_t es = clo_type_sizeof(scanner->elem_type), ss = clo_type_sizeof(scanner->sum_type);
			bool ready = true;
			if (!scanner->s_in) {
				ready = cudaStreamCreateWithFlags(&scanner->s_in, cudaStreamNonBlocking) == cudaSuccess &&
					cudaStreamCreateWithFlags(&scanner->s_out, cudaStreamNonBlocking) == cudaSuccess &&
					cudaMalloc(&scanner->d_chunk_carry, 16) == cudaSuccess;
				for (int i = 0; ready && i < 4; ++i)
					ready = cudaEventCreateWithFlags(&scanner->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
						cudaEventCreateWithFlags(&scanner->ev_sc[i], cudaEventDisableTiming) == cudaSuccess;
			}
			if (clo_cuda_failed(ready ? cudaSuccess : cudaErrorMemoryAllocation, err, "scan host pipeline")) goto done;
			cudaStream_t se = cq_exec->stream;
			const int sms = clo_sm_count(cq_exec->ctx->dev.ordinal);
			/* the copy streams start after whatever the caller already queued */
			cudaEventRecord(scanner->ev_sc[3], se);
			cudaStreamWaitEvent(scanner->s_in, scanner->ev_sc[3], 0);
			cudaStreamWaitEvent(scanner->s_out, scanner->ev_sc[3], 0);
			cudaError_t rc = cudaSuccess;
			size_t idx = 0;
			for (size_t off = 0; off < numel && rc == cudaSuccess; off += chunk, ++idx) {
				const size_t cnt = numel - off < chunk ? numel - off : chunk;
				const int k = (int) (idx % 3);
				const char* din = (const char*) scanner->host_in.ptr + off * es;
				char* dout = (char*) scanner->host_out.ptr + off * ss;
				rc = cudaMemcpyAsync((void*) din, (const char*) data_in + off * es, cnt * es, cudaMemcpyHostToDevice, scanner->s_in);
				if (rc == cudaSuccess) rc = cudaEventRecord(scanner->ev_in[k], scanner->s_in);
				if (rc == cudaSuccess) rc = cudaStreamWaitEvent(se, scanner->ev_in[k], 0);
				if (rc == cudaSuccess) rc = scanner->fn(scanner->st, din, dout, cnt, off ? scanner->d_chunk_carry : NULL, sms, se);
				if (rc == cudaSuccess && off + cnt < numel) rc = scanner->cfn(din, dout, cnt, scanner->d_chunk_carry, se);
				if (rc == cudaSuccess) rc = cudaEventRecord(scanner->ev_sc[k], se);
				if (rc == cudaSuccess) rc = cudaStreamWaitEvent(scanner->s_out, scanner->ev_sc[k], 0);
				if (rc == cudaSuccess) rc = cudaMemcpyAsync((char*) data_out + off * ss, dout, cnt * ss, cudaMemcpyDeviceToHost, scanner->s_out);
				/* an event slot is reused three chunks later: its waiters have been queued long before */
			}
			if (rc == cudaSuccess) rc = cudaStreamSynchronize(scanner->s_out);       /* the one blocking point */
			if (rc == cudaSuccess) rc = cudaStreamSynchronize(se);
			if (clo_cuda_failed(rc, err, "scan host pipeline")) goto done;
			if (scan_check_error_flag(scanner, se, err)) goto done;
			ok = CL_TRUE;
			goto done;
		}
		CCLEvent* evt = ccl_buffer_enqueue_write(in_dev, cq_comm, CL_FALSE, 0, in_size, data_in, NULL, err);
		if (!evt) goto done;
		if (!one_queue && !ccl_queue_finish(cq_comm, err)) goto done;
		evt = scanner->impl_def.scan_with_device_data(scanner, cq_exec, cq_comm, in_dev, out_dev, numel, lws_max, err);
		if (!evt) goto done;
		evt = ccl_buffer_enqueue_read(out_dev, cq_comm, CL_FALSE, 0, out_size, data_out,
			!one_queue ? ccl_ewl(&ewl, evt, NULL) : NULL, err);
		if (!evt) goto done;
		if (!ccl_queue_finish(cq_comm, err)) goto done;                 /* the one blocking point */
		if (!one_queue && !ccl_queue_finish(cq_exec, err)) goto done;
		if (scan_check_error_flag(scanner, cq_exec->stream, err)) goto done;
		ok = CL_TRUE;
	}
done:
	ccl_event_wait_list_clear(&ewl);
	if (in_dev) ccl_buffer_destroy(in_dev);
	if (out_dev) ccl_buffer_destroy(out_dev);
	if (own_queue) ccl_queue_destroy(own_queue);
	return ok;
}

extern "C" CCLContext* clo_scan_get_context(CloScan* s) { return s ? s->ctx : NULL; }
extern "C" CCLProgram* clo_scan_get_program(CloScan* s) { return s ? s->prg : NULL; }
extern "C" CloType clo_scan_get_elem_type(CloScan* s) { return s ? s->elem_type : (CloType) -1; }
extern "C" size_t clo_scan_get_element_size(CloScan* s) { return s ? clo_type_sizeof(s->elem_type) : 0; }
extern "C" CloType clo_scan_get_sum_type(CloScan* s) { return s ? s->sum_type : (CloType) -1; }
extern "C" size_t clo_scan_get_sum_size(CloScan* s) { return s ? clo_type_sizeof(s->sum_type) : 0; }
extern "C" void* clo_scan_get_data(CloScan* s) { return s ? s->data : NULL; }
extern "C" void clo_scan_set_data(CloScan* s, void* data) { if (s) s->data = data; }
extern "C" cl_uint clo_scan_get_num_kernels(CloScan* s, GError** err) { return s ? s->impl_def.get_num_kernels(s, err) : 0; }
extern "C" const char* clo_scan_get_kernel_name(CloScan* s, cl_uint i, GError** err) {
	return s ? s->impl_def.get_kernel_name(s, i, err) : NULL;
}
extern "C" size_t clo_scan_get_localmem_usage(CloScan* s, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	return s ? s->impl_def.get_localmem_usage(s, i, lws_max, numel, err) : 0;
}
