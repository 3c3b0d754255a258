/*
 * partition.cu -- the sample sort's stable multi-way partition (the
 * reference is single-device, /root/reference/src/cl_ops/sort/clo_sort_abstract.c:335, so
 * there is nothing to restate: the contract is the one in include/cl_ops/clo_b200.h).
 *
 * With at most 16 buckets there is no need for the radix machinery (tile staging, look-back):
 *   K1 count    every WARP owns one contiguous chunk of the input and counts its keys per
 *               bucket (one aggregated shared-memory RED per 32 keys)
 *   K2 offsets  one CTA per bucket scans the per-warp counts -> each warp's first output
 *               slot in every bucket, and the bucket totals (counts_out)
 *   K3 scatter  every warp walks its chunk again; a key's slot is
 *               bucket start + warp offset + running count + rank among the lower lanes,
 *               the rank coming from one ballot per bucket.  Lanes of a bucket write
 *               consecutive addresses, instruction after instruction, so the partial
 *               sectors merge in L2.
 * HBM traffic: 2 reads + 1 write of the keys (12 B per u32 key), no atomics on global memory.
 * Stable: chunks are in input order, a warp walks its chunk in order, ranks follow lane order.
 */
#include "clo_internal.h"
#include "device_utils.cuh"
#include "sort_common.h"

#include <type_traits>

using namespace clo;

namespace {

const int PT_THREADS = 256;
const int PT_WARPS = PT_THREADS / 32;
const int PT_MAXP = 16;
const int PT_U = 8;                       /* independent 128-byte loads in flight per warp */

/* NS = number of splitter slots compiled in (1, 3, 7 or 15); slots past the real splitters
 * hold (max key, max index), which no element reaches, so they never count */
template <typename ElemT, int NS> struct PtSplitters {
	ElemT key[NS];
	u64 idx[NS];
	u64 gidx0;                            /* global index of element 0 */
};

/* the splitters are tiny and device resident: every thread pulls them into registers */
template <typename ElemT, int NS>
__device__ __forceinline__ void pt_load_splitters(PtSplitters<ElemT, NS>& sp, const ElemT* __restrict__ sk,
		const u64* __restrict__ si, u32 count, u64 gidx0) {
	sp.gidx0 = gidx0;
#pragma unroll
	for (int s = 0; s < NS; ++s) {
		sp.key[s] = s < (int) count ? sk[s] : (ElemT) ~(ElemT) 0;
		sp.idx[s] = s < (int) count ? si[s] : ~0ull;
	}
}

/* compile-time-indexed element of a register array chosen by a run-time index (select chain) */
template <typename T, int N>
__device__ __forceinline__ T pt_pick(const T (&a)[N], u32 idx, int stride, int first) {
	T v = a[first];
#pragma unroll
	for (int j = first + stride; j < N; j += stride) if (idx >= (u32) (j - first)) v = a[j];
	return v;
}

/* bucket = number of splitters (key, index) <= (k, g).  The splitter keys are sorted, so the
 * number of keys < k is a branch-free binary search over the NS = 2^m - 1 register-resident
 * splitters; the search always compares k with the first splitter key >= k, so "some pivot
 * equalled k" detects exactly the keys whose bucket depends on the 64-bit index tie-break,
 * which then (and only then, for the whole warp) runs the linear rule. */
template <typename ElemT, int NS>
__device__ __forceinline__ u32 pt_bucket(ElemT k, u64 g, const PtSplitters<ElemT, NS>& sp) {
	u32 b = 0;
	bool tie = false;
#pragma unroll
	for (int step = (NS + 1) / 2; step >= 1; step >>= 1) {
		/* candidates at this level: indices step-1, step-1 + 2*step, ... ; b is a multiple of 2*step */
		const ElemT pivot = pt_pick<ElemT, NS>(sp.key, b, 2 * step, step - 1);
		tie |= pivot == k;
		if (pivot < k) b += (u32) step;
	}
	if (__any_sync(__activemask(), tie)) {
		b = 0;
#pragma unroll
		for (int s = 0; s < NS; ++s) b += (sp.key[s] < k || (sp.key[s] == k && sp.idx[s] <= g)) ? 1u : 0u;
	}
	return b;
}

/* bucket of one element when nobody else in the warp has to agree on the path (the count kernel):
 * the tie-break is an ordinary divergent branch */
template <typename ElemT, int NS>
__device__ __forceinline__ u32 pt_bucket_lane(ElemT k, u64 g, const PtSplitters<ElemT, NS>& sp) {
	u32 b = 0;
	bool tie = false;
#pragma unroll
	for (int step = (NS + 1) / 2; step >= 1; step >>= 1) {
		const ElemT pivot = pt_pick<ElemT, NS>(sp.key, b, 2 * step, step - 1);
		tie |= pivot == k;
		if (pivot < k) b += (u32) step;
	}
	if (tie) {
		b = 0;
#pragma unroll
		for (int s = 0; s < NS; ++s) b += (sp.key[s] < k || (sp.key[s] == k && sp.idx[s] <= g)) ? 1u : 0u;
	}
	return b;
}

/* the same bucket, with the sorted splitter keys in SHARED memory: three broadcast-friendly loads
 * replace the register select chains of pt_pick (which made the 8-way count ALU bound) */
template <typename ElemT, int NS>
__device__ __forceinline__ u32 pt_bucket_smem(ElemT k, u64 g, const ElemT* __restrict__ s_sk, const PtSplitters<ElemT, NS>& sp) {
	u32 b = 0;
	bool tie = false;
#pragma unroll
	for (int step = (NS + 1) / 2; step >= 1; step >>= 1) {
		const ElemT pivot = s_sk[b + (u32) step - 1u];
		tie |= pivot == k;
		b += pivot < k ? (u32) step : 0u;
	}
	if (tie) {
		b = 0;
#pragma unroll
		for (int s = 0; s < NS; ++s) b += (sp.key[s] < k || (sp.key[s] == k && sp.idx[s] <= g)) ? 1u : 0u;
	}
	return b;
}

/* Bucket sizes per scatter chunk (chunk w = [w * chunk, min(n, (w + 1) * chunk)) is what warp w of
 * the scatter kernel will move).  Round 2: counting streams through the keys like any grid-stride
 * kernel -- sub-blocks of PT_SUB keys, handed to the warps round-robin, so the chip reads one
 * moving window instead of one far-apart stream per chunk (3.6 -> 5 TB/s) -- with no shared-memory
 * atomics and no warp votes: every lane counts its keys in byte-wide fields of one (<= 8 buckets)
 * or two 64-bit registers, the warp adds the lanes up per sub-block, and lane q adds bucket q's
 * count to the chunk the sub-block belongs to (chunk is a multiple of PT_SUB; counts zeroed by
 * the launcher). */
const int PT_SUB = 1024;
template <typename ElemT, int NS>
__global__ void __launch_bounds__(PT_THREADS)
clo_partition_count(const ElemT* __restrict__ in, size_t n, size_t chunk, u32 total_warps,
		const ElemT* __restrict__ sk, const u64* __restrict__ si, u32 nsplit, u64 gidx0,
		u32* __restrict__ counts /* [PT_MAXP][total_warps] */) {
	constexpr int NB = NS + 1;                          /* buckets compiled in: 2, 4, 8 or 16 */
	constexpr int EPV = 16 / (int) sizeof(ElemT);
	constexpr int VPL = PT_SUB / 32 / EPV;              /* 16-byte vectors per lane and sub-block */
	PtSplitters<ElemT, NS> sp;
	pt_load_splitters<ElemT, NS>(sp, sk, si, nsplit, gidx0);
	constexpr bool SMEM_SEARCH = NS >= 3;
	__shared__ ElemT s_sk[NS + 1];
	if (SMEM_SEARCH) {
		if (threadIdx.x < NS) s_sk[threadIdx.x] = threadIdx.x < nsplit ? sk[threadIdx.x] : (ElemT) ~(ElemT) 0;
		__syncthreads();
	}
	const int lane = threadIdx.x & 31;
	const size_t gw = (size_t) blockIdx.x * PT_WARPS + (threadIdx.x >> 5);
	const size_t nwarps = (size_t) gridDim.x * PT_WARPS;
	const size_t nsub = (n + PT_SUB - 1) / PT_SUB;
	const bool vec_ok = (reinterpret_cast<uintptr_t>(in) % 16) == 0;
	for (size_t sb = gw; sb < nsub; sb += nwarps) {
		const size_t lo = sb * PT_SUB;
		const size_t hi = lo + PT_SUB < n ? lo + PT_SUB : n;
		const bool full = vec_ok && lo + PT_SUB <= n;
		u64 acc0 = 0, acc1 = 0;
		auto count_one = [&](ElemT k, size_t i) {
			const u32 bkt = SMEM_SEARCH ? pt_bucket_smem<ElemT, NS>(k, sp.gidx0 + i, s_sk, sp) : pt_bucket_lane<ElemT, NS>(k, sp.gidx0 + i, sp);
			const u64 inc = 1ull << ((bkt & 7u) << 3);
			if (NB <= 8) acc0 += inc;
			else { acc0 += bkt < 8u ? inc : 0ull; acc1 += bkt < 8u ? 0ull : inc; }
		};
		if (full) {
			constexpr int VB = VPL > 8 ? 8 : VPL;           /* vectors in flight per lane (32 registers of keys) */
#pragma unroll
			for (int u0 = 0; u0 < VPL; u0 += VB) {
				ElemT k[VB][EPV];
#pragma unroll
				for (int u = 0; u < VB; ++u) load_vec_cs<ElemT, EPV>(in + lo + ((size_t) (u0 + u) * 32 + lane) * EPV, k[u]);
#pragma unroll
				for (int u = 0; u < VB; ++u)
#pragma unroll
					for (int c = 0; c < EPV; ++c) count_one(k[u][c], lo + ((size_t) (u0 + u) * 32 + lane) * EPV + c);
			}
		} else {
			for (size_t i = lo + lane; i < hi; i += 32) count_one(__ldcs(in + i), i);
		}
		const size_t cw = lo / chunk;                       /* the scatter warp this sub-block belongs to */
		u32 mine = 0;
#pragma unroll
		for (int q = 0; q < NB; ++q) {
			const u32 t = warp_reduce_sum<u32>((u32) (((q < 8 ? acc0 : acc1) >> ((q & 7) << 3)) & 0xffull));
			if (lane == q) mine = t;
		}
		if (lane < NB && mine) atomicAdd(&counts[(size_t) lane * total_warps + cw], mine);
	}
}

/* block q: exclusive scan of counts[q][*] in place (-> warp offsets inside bucket q) and the
 * bucket total */
__global__ void __launch_bounds__(1024)
clo_partition_offsets(u32* __restrict__ counts, u32 total_warps, u64* __restrict__ totals,
		u64* __restrict__ counts_out, u32 nparts) {
	__shared__ u64 s_w[32];
	__shared__ u64 s_run;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	u32* row = counts + (size_t) blockIdx.x * total_warps;
	if (t == 0) s_run = 0;
	__syncthreads();
	for (u32 base = 0; base < total_warps; base += 1024) {
		const u32 i = base + t;
		const u64 c = i < total_warps ? row[i] : 0;
		const u64 incl = warp_inclusive_scan<u64>(c, lane);
		if (lane == 31) s_w[warp] = incl;
		__syncthreads();
		u64 off = s_run;
		for (int v = 0; v < warp; ++v) off += s_w[v];
		/* offsets inside a bucket fit 32 bits (n < 2^32 is checked by the host side) */
		if (i < total_warps) row[i] = (u32) (off + incl - c);
		__syncthreads();
		if (t == 1023) s_run = off + incl;
		__syncthreads();
	}
	if (t == 0) {
		totals[blockIdx.x] = s_run;
		if (counts_out && blockIdx.x < nparts) counts_out[blockIdx.x] = s_run;
	}
}

/* local mode: every bucket goes to the same output array, bucket after bucket */
__global__ void clo_partition_local_slots(const u64* __restrict__ totals, u64* __restrict__ first_slot,
		void** __restrict__ dests, void** __restrict__ vdests, void* out, void* vout) {
	const int q = threadIdx.x;
	if (q < PT_MAXP) {
		u64 run = 0;
		for (int p = 0; p < q; ++p) run += totals[p];
		first_slot[q] = run;
		dests[q] = out;
		vdests[q] = vout;
	}
}

/* Bucket q of this rank goes to dests[q] (possibly PEER memory, written over NVLink) starting
 * at element first_slot[q].  Keys are not written as they come (a warp instruction would give
 * each bucket a 16-64 byte fragment, which NVLink and the L2 handle badly): every warp keeps a
 * 64-element ring per bucket in shared memory and writes a bucket only in whole, 128-byte
 * ALIGNED lines of 32 elements (the first write of a bucket is short so that the following ones
 * are aligned; the last one drains the ring).  Lane q carries the state of bucket q.
 * *ok == 0 (a receive buffer would overflow) turns the kernel into a no-op. */
template <typename ElemT, bool HAS_VAL, int NS, int SW /* warps per CTA */>
__global__ void __launch_bounds__(SW * 32)
clo_partition_scatter(const ElemT* __restrict__ in, const u32* __restrict__ vin, size_t n, size_t chunk, u32 total_warps,
		const ElemT* __restrict__ sk, const u64* __restrict__ si, u32 nsplit, u64 gidx0,
		const u32* __restrict__ offsets /* [PT_MAXP][total_warps] */, const u64* __restrict__ first_slot,
		ElemT* const* __restrict__ dests, u32* const* __restrict__ vdests, const int* __restrict__ ok, u32 nparts) {
	constexpr int NB = NS + 1;
	__shared__ ElemT s_key[SW][NB][64];
	__shared__ u32 s_val[HAS_VAL ? SW : 1][HAS_VAL ? NB : 1][HAS_VAL ? 64 : 1];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 w = blockIdx.x * SW + warp;
	if (w >= total_warps) return;
	if (ok && *ok == 0) return;
	PtSplitters<ElemT, NS> sp;
	pt_load_splitters<ElemT, NS>(sp, sk, si, nsplit, gidx0);
	/* splitter keys also in shared memory: the bucket search is three loads instead of select chains */
	constexpr size_t RING_BYTES = (size_t) SW * NB * 64 * (sizeof(ElemT) + (HAS_VAL ? 4 : 0));
	constexpr bool SMEM_SEARCH = NS >= 3 && RING_BYTES + (NS + 1) * sizeof(ElemT) <= 48 * 1024;
	__shared__ ElemT s_sk[SMEM_SEARCH ? NS + 1 : 1];
	if (SMEM_SEARCH) {
		/* written by every warp with the same values: no CTA barrier (some warps have returned) */
		if (lane < NS) s_sk[lane] = (u32) lane < nsplit ? sk[lane] : (ElemT) ~(ElemT) 0;
		__syncwarp();
	}
	ElemT* kp = nullptr;          /* next element of bucket `lane` to be written */
	u32* vp = nullptr;
	u32 head = 0, cnt = 0;        /* ring of bucket `lane`: first pending slot, pending elements */
	if (lane < (int) nparts) {
		const u64 slot = first_slot[lane] + offsets[(size_t) lane * total_warps + w];
		kp = dests[lane] + slot;
		if (HAS_VAL) vp = vdests[lane] + slot;
	}
	u32 thr = 32u - (u32) ((reinterpret_cast<uintptr_t>(kp) & 127) / sizeof(ElemT));
	ElemT (*ring)[64] = s_key[warp];
	const size_t lo = (size_t) w * chunk;
	const size_t hi = lo + chunk < n ? lo + chunk : n;
	const u32 lt = lanemask_lt();
	/* write f pending elements of bucket q (state in lane q) */
	auto flush = [&](int q, u32 f) {
		const u32 h = __shfl_sync(0xffffffffu, head, q);
		ElemT* dk = reinterpret_cast<ElemT*>(__shfl_sync(0xffffffffu, (u64) reinterpret_cast<uintptr_t>(kp), q));
		if ((u32) lane < f) dk[lane] = ring[q][(h + lane) & 63];
		if (HAS_VAL) {
			u32* dv = reinterpret_cast<u32*>(__shfl_sync(0xffffffffu, (u64) reinterpret_cast<uintptr_t>(vp), q));
			if ((u32) lane < f) dv[lane] = s_val[HAS_VAL ? warp : 0][HAS_VAL ? q : 0][HAS_VAL ? ((h + lane) & 63) : 0];
		}
		if (lane == q) { head = (head + f) & 63; cnt -= f; kp += f; if (HAS_VAL) vp += f; thr = 32u; }
	};
	/* one step = PT_U rows of 32 keys; FULL steps (all but the last of the array) skip every bounds check */
	auto step_rows = [&](size_t base, auto full_tag) {
		constexpr bool FULL = decltype(full_tag)::value;
		ElemT k[PT_U];
		u32 v[HAS_VAL ? PT_U : 1];
#pragma unroll
		for (int u = 0; u < PT_U; ++u) {
			const size_t i = base + u * 32 + lane;
			k[u] = (FULL || i < hi) ? __ldcs(in + i) : ElemT(0);
			if (HAS_VAL) v[u] = (FULL || i < hi) ? __ldcs(vin + i) : 0u;
		}
#pragma unroll
		for (int u = 0; u < PT_U; ++u) {
			const size_t i = base + u * 32 + lane;
			const bool valid = FULL || i < hi;
			u32 b;
			if (SMEM_SEARCH) {
				b = 0;
				bool tie = false;
#pragma unroll
				for (int step = (NS + 1) / 2; step >= 1; step >>= 1) {
					const ElemT pivot = s_sk[b + (u32) step - 1u];
					tie |= pivot == k[u];
					b += pivot < k[u] ? (u32) step : 0u;
				}
				if (__any_sync(0xffffffffu, tie)) b = pt_bucket<ElemT, NS>(k[u], sp.gidx0 + i, sp);     /* rare: the exact rule, whole warp */
			} else {
				b = pt_bucket<ElemT, NS>(k[u], sp.gidx0 + i, sp);
			}
			if (!valid) b = 0xffffffffu;
			/* lanes of my bucket / lanes of bucket `lane`: one ballot per bucket BIT */
			u32 mine = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, valid), forq = mine;
#pragma unroll
			for (int j = 0; (1 << j) < NB; ++j) {
				const u32 bj = __ballot_sync(0xffffffffu, (b >> j) & 1u);
				mine &= ((b >> j) & 1u) ? bj : ~bj;
				forq &= ((lane >> j) & 1) ? bj : ~bj;
			}
			const u32 add = lane < NB ? __popc(forq) : 0u;
			/* append to the ring of my bucket */
			const u32 tail = __shfl_sync(0xffffffffu, head + cnt, valid ? (int) b : 0);
			if (valid) {
				const u32 slot = (tail + __popc(mine & lt)) & 63;
				ring[b][slot] = k[u];
				if (HAS_VAL) s_val[HAS_VAL ? warp : 0][HAS_VAL ? b : 0][HAS_VAL ? slot : 0] = v[u];
			}
			cnt += add;
			__syncwarp();
			/* buckets that can fill a line up to the next 128-byte boundary (thr: elements up to it;
			 * a flush always ends on the boundary, so it is 32 from the bucket's first flush on) */
			for (;;) {
				u32 need = __ballot_sync(0xffffffffu, lane < NB && cnt >= thr);
				if (!need) break;
				while (need) {
					const int q = __ffs(need) - 1;
					need &= need - 1;
					flush(q, __shfl_sync(0xffffffffu, thr, q));
				}
				__syncwarp();
			}
		}
	};
	for (size_t base = lo; base < hi; base += 32 * PT_U) {
		if (base + 32 * PT_U <= hi) step_rows(base, std::true_type{});
		else step_rows(base, std::false_type{});
	}
	/* drain */
	for (int q = 0; q < NB; ++q) {
		u32 left = __shfl_sync(0xffffffffu, cnt, q);
		while (left) {
			const u32 f = left < 32u ? left : 32u;
			flush(q, f);
			left -= f;
		}
	}
}

struct PtPlan { size_t chunk; u32 total_warps, grid; u32* counts; u64* totals; u64* first_slot; void** dests; void** vdests; };

cudaError_t pt_plan(CloScratch& work, size_t n, int sm_count, PtPlan& pl) {
	const u32 ctas = (u32) sm_count * 8;
	u32 total_warps = ctas * PT_WARPS;
	size_t chunk = (n + total_warps - 1) / total_warps;
	const size_t gran = PT_SUB;                       /* a multiple of 32 * PT_U; the count kernel's sub-block */
	chunk = (chunk + gran - 1) / gran * gran;
	if (chunk == 0) chunk = gran;
	/* the warp count is a function of sm_count only, so the count and scatter stages of one
	 * partition (and the scratch they share) always agree; idle warps have empty chunks */
	pl.chunk = chunk;
	pl.total_warps = total_warps;
	pl.grid = ctas;
	const size_t cnt_bytes = ((size_t) PT_MAXP * total_warps * sizeof(u32) + 255) / 256 * 256;
	cudaError_t e = work.reserve(cnt_bytes + 4 * 256);
	if (e != cudaSuccess) return e;
	char* base = (char*) work.ptr;
	pl.counts = (u32*) base;
	pl.totals = (u64*) (base + cnt_bytes);
	pl.first_slot = (u64*) (base + cnt_bytes + 256);
	pl.dests = (void**) (base + cnt_bytes + 512);
	pl.vdests = (void**) (base + cnt_bytes + 768);
	return cudaSuccess;
}

template <typename ElemT>
cudaError_t pt_count_typed(CloScratch& work, const ElemT* in, size_t n, u64 gidx0, const void* sk, const u64* si,
		u32 nparts, u64* counts_out, int sm_count, cudaStream_t stream) {
	PtPlan pl;
	cudaError_t e = pt_plan(work, n, sm_count, pl);
	if (e != cudaSuccess) return e;
	const ElemT* dsk = (const ElemT*) sk;
	const u32 nsplit = nparts - 1;
	auto run = [&](auto ns_tag) {
		constexpr int NS = decltype(ns_tag)::value;
		clo_partition_count<ElemT, NS><<<pl.grid, PT_THREADS, 0, stream>>>(in, n, pl.chunk, pl.total_warps, dsk, si, nsplit, gidx0, pl.counts);
	};
	if ((e = cudaMemsetAsync(pl.counts, 0, (size_t) PT_MAXP * pl.total_warps * sizeof(u32), stream)) != cudaSuccess) return e;
	if (nsplit <= 1) run(std::integral_constant<int, 1>{});
	else if (nsplit <= 3) run(std::integral_constant<int, 3>{});
	else if (nsplit <= 7) run(std::integral_constant<int, 7>{});
	else run(std::integral_constant<int, 15>{});
	clo_partition_offsets<<<PT_MAXP, 1024, 0, stream>>>(pl.counts, pl.total_warps, pl.totals, counts_out, nparts);
	CLO_COUNT_LAUNCH(2);
	return cudaGetLastError();
}

template <typename ElemT>
cudaError_t pt_scatter_typed(CloScratch& work, const ElemT* in, const u32* vin, size_t n, u64 gidx0, const void* sk,
		const u64* si, u32 nparts, const u64* first_slot, void* const* dests, void* const* vdests, const int* ok,
		int sm_count, cudaStream_t stream) {
	PtPlan pl;
	cudaError_t e = pt_plan(work, n, sm_count, pl);
	if (e != cudaSuccess) return e;
	const ElemT* dsk = (const ElemT*) sk;
	const u32 nsplit = nparts - 1;
	auto run = [&](auto ns_tag) {
		constexpr int NS = decltype(ns_tag)::value;
		/* the rings of 16 buckets need 48 KB for 4 warps: smaller CTAs there */
		constexpr int SW = NS > 7 ? 4 : PT_WARPS;
		const u32 grid = (pl.total_warps + SW - 1) / SW;
		if (vin)
			clo_partition_scatter<ElemT, true, NS, SW><<<grid, SW * 32, 0, stream>>>(in, vin, n, pl.chunk, pl.total_warps, dsk, si, nsplit, gidx0,
				pl.counts, first_slot, (ElemT* const*) dests, (u32* const*) vdests, ok, nparts);
		else
			clo_partition_scatter<ElemT, false, NS, SW><<<grid, SW * 32, 0, stream>>>(in, vin, n, pl.chunk, pl.total_warps, dsk, si, nsplit, gidx0,
				pl.counts, first_slot, (ElemT* const*) dests, (u32* const*) vdests, ok, nparts);
	};
	if (nsplit <= 1) run(std::integral_constant<int, 1>{});
	else if (nsplit <= 3) run(std::integral_constant<int, 3>{});
	else if (nsplit <= 7) run(std::integral_constant<int, 7>{});
	else run(std::integral_constant<int, 15>{});
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

bool pt_args_ok(size_t elem_size, size_t n, uint32_t nparts, const char** err_msg) {
	if (nparts < 1 || nparts > (uint32_t) PT_MAXP) { if (err_msg) *err_msg = "partition: nparts must be in [1,16]"; return false; }
	if (n >= (1ull << 32)) { if (err_msg) *err_msg = "partition: too many elements"; return false; }
	if (elem_size != 4 && elem_size != 8) { if (err_msg) *err_msg = "partition: keys must be 4 or 8 bytes"; return false; }
	return true;
}

} // namespace

/* stage 1: bucket sizes of this rank -> counts_out[nparts] (device); keeps the per-warp
 * offsets in `work` for stage 2 (same n, same device) */
cudaError_t clo_partition_count_stage(CloScratch& work, size_t elem_size, const void* keys_in, size_t n, uint64_t gidx0,
		const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out, int sm_count,
		cudaStream_t stream, const char** err_msg) {
	if (!pt_args_ok(elem_size, n, nparts, err_msg)) return cudaErrorInvalidValue;
	if (elem_size == 4)
		return pt_count_typed<u32>(work, (const u32*) keys_in, n, gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, sm_count, stream);
	return pt_count_typed<u64>(work, (const u64*) keys_in, n, gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, sm_count, stream);
}

/* stage 2: scatter bucket q to dests[q] + first_slot[q] (device arrays of nparts entries;
 * the destinations may be peer memory) */
cudaError_t clo_partition_scatter_stage(CloScratch& work, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		size_t n, uint64_t gidx0, const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts,
		const uint64_t* first_slot, void* const* dests, void* const* vdests, const int* ok, int sm_count,
		cudaStream_t stream, const char** err_msg) {
	if (!pt_args_ok(elem_size, n, nparts, err_msg)) return cudaErrorInvalidValue;
	if (elem_size == 4)
		return pt_scatter_typed<u32>(work, (const u32*) keys_in, payload_in, n, gidx0, splitter_keys, (const u64*) splitter_idx, nparts,
			(const u64*) first_slot, dests, vdests, ok, sm_count, stream);
	return pt_scatter_typed<u64>(work, (const u64*) keys_in, payload_in, n, gidx0, splitter_keys, (const u64*) splitter_idx, nparts,
		(const u64*) first_slot, dests, vdests, ok, sm_count, stream);
}

cudaError_t clo_partition_v2(CloScratch& work, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		void* keys_out, uint32_t* payload_out, size_t n, uint64_t gidx0, const void* splitter_keys,
		const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out, int sm_count, cudaStream_t stream,
		const char** err_msg) {
	cudaError_t e = clo_partition_count_stage(work, elem_size, keys_in, n, gidx0, splitter_keys, splitter_idx, nparts,
		counts_out, sm_count, stream, err_msg);
	if (e != cudaSuccess) return e;
	PtPlan pl;
	if ((e = pt_plan(work, n, sm_count, pl)) != cudaSuccess) return e;
	clo_partition_local_slots<<<1, 32, 0, stream>>>(pl.totals, pl.first_slot, pl.dests, pl.vdests, keys_out, payload_out);
	CLO_COUNT_LAUNCH(1);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	return clo_partition_scatter_stage(work, elem_size, keys_in, payload_in, n, gidx0, splitter_keys, splitter_idx, nparts,
		(const uint64_t*) pl.first_slot, pl.dests, pl.vdests, nullptr, sm_count, stream, err_msg);
}
