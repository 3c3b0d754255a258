/*
 * scan_tma.cuh -- the persistent single-pass scan for element and sum types of the SAME size,
 * rebuilt around the Blackwell copy engine.  Included by scan.cu after scan_pp.cuh (shares its
 * AGG/PREF protocol and the propagator CTA).  Replaces clo_scan_blelloch.cl:49-211.
 *
 * What ncu said about clo_scan_pp (profiles/r02_scan_pp_f32_ncu_summary.txt): 25 % occupancy,
 * issue slots 32 % busy, a warp issues once every 12.5 cycles (barrier 3.9, short scoreboard
 * 3.7, wait 1.9) and runs ~340 instructions per 4096-element tile -- the kernel is bound by
 * the latency of its own instruction chain, not by HBM: four warp scans (7 shuffles each) per
 * thread and tile because a lane's four vectors are 512 bytes apart, 8 per-thread cp.async /
 * stores with their address arithmetic, three CTA barriers.  Here
 *   - ONE thread moves the tiles: cp.async.bulk.tensor (TMA) global -> shared with an mbarrier
 *     per ring slot, and shared -> global for the results (bulk groups), so no thread computes
 *     a global address;
 *   - the tensor maps use SWIZZLE_128B (16-byte chunk index ^= 128-byte row index & 7), which
 *     lets every thread own 64 CONTIGUOUS bytes of the tile and still read and write them with
 *     conflict-free 16-byte shared-memory accesses (8 consecutive lanes hit 8 different chunk
 *     columns): one warp scan per tile instead of four, done when the tile is REDUCED -- its
 *     per-thread exclusive sums wait in shared memory next to the tile;
 *   - the scan phase is shuffle free: base = PREF[tile] + warps below + lanes below, a serial
 *     pass over the thread's 16 (8) elements, written back IN PLACE, then one TMA store;
 *   - one CTA barrier per tile.
 * HBM traffic is unchanged: every element is read once and written once.
 */
#ifndef CLO_SCAN_TMA_CUH
#define CLO_SCAN_TMA_CUH

#include <cuda.h>

__device__ __forceinline__ unsigned stm_smem(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void stm_mbar_init(u64* bar, unsigned count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(stm_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void stm_mbar_expect_tx(u64* bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(stm_smem(bar)), "r"(bytes) : "memory");
}
/* bounded: a copy that never lands (a bad descriptor, a fault) must not hang the GPU; false = timed out */
__device__ __forceinline__ bool stm_mbar_wait(u64* bar, unsigned parity) {
	for (unsigned tries = 0; tries < (1u << 24); ++tries) {
		unsigned ok;
		asm volatile(
			"{\n"
			".reg .pred p;\n"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			"selp.u32 %0, 1, 0, p;\n"
			"}\n" : "=r"(ok) : "r"(stm_smem(bar)), "r"(parity) : "memory");
		if (ok) return true;
	}
	return false;
}
__device__ __forceinline__ void stm_tma_load(void* smem_dst, const CUtensorMap* tm, int c0, int c1, u64* bar) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		:: "r"(stm_smem(smem_dst)), "l"(reinterpret_cast<u64>(tm)), "r"(c0), "r"(c1), "r"(stm_smem(bar)) : "memory");
}
__device__ __forceinline__ void stm_tma_store(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
		:: "l"(reinterpret_cast<u64>(tm)), "r"(stm_smem(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void stm_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void stm_bulk_wait_read() {
	asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
__device__ __forceinline__ void stm_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

/* SLACK: bulk-store groups that may still be reading shared memory when the next load is
 * issued (1 needs one more ring slot than 0) */
template <int THREADS, int AHEAD_, int LAG_, int SLACK>
struct StmShape {
	static constexpr int S = AHEAD_ + LAG_ + 1 + SLACK;
	static constexpr int TILE_BYTES = THREADS * 64;
	static constexpr int ROWS = TILE_BYTES / 128;               /* 128-byte rows per tile (box height) */
};

template <typename ElemT, typename SumT, int THREADS, int AHEAD_, int LAG_, int SLACK>
__global__ void __launch_bounds__(THREADS, 2)
clo_scan_tma(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
		u32 num_tiles, u64* __restrict__ agg, u64* __restrict__ pref, u32* __restrict__ ticket, u32 epoch,
		const SumT* __restrict__ carry_in, int* __restrict__ err_flag, int flags) {
	typedef typename AccOf<SumT>::type AccT;
	typedef AccWords<AccT> AW;
	typedef typename std::conditional<std::is_same<SumT, float>::value, float, AccT>::type IntraT;
	static_assert(sizeof(ElemT) == sizeof(SumT), "results replace the elements in the ring");
	static_assert(AHEAD_ >= 1, "a slot is loaded at least one iteration before it is reduced");
	typedef StmShape<THREADS, AHEAD_, LAG_, SLACK> Shape;
	constexpr int S = Shape::S;
	constexpr int NT = S + 2;
	constexpr int WARPS = THREADS / 32;
	constexpr int EPV = 16 / (int) sizeof(ElemT);            /* elements per 16-byte chunk */
	constexpr int TILE_BYTES = Shape::TILE_BYTES;
	constexpr int ROWS = Shape::ROWS;
	constexpr bool FLOAT32 = std::is_same<SumT, float>::value;

	const AccT carry = carry_in ? to_acc<SumT, SumT, AccT>(*carry_in) : AccT(0);
	if (blockIdx.x == 0) {
		if (flags & 1) spp_propagate<AccT, THREADS>(agg, pref, num_tiles, epoch, carry, err_flag);
		else if (flags & 2) spp_propagate_chain<AccT, THREADS, 2>(agg, pref, num_tiles, epoch, carry, err_flag);
		else spp_propagate_chain<AccT, THREADS, 1>(agg, pref, num_tiles, epoch, carry, err_flag);
		return;
	}

	extern __shared__ unsigned char stm_raw[];
	/* SWIZZLE_128B repeats every 1024 bytes: the ring starts on such a boundary */
	unsigned char* ring = stm_raw + ((1024u - (stm_smem(stm_raw) & 1023u)) & 1023u);   /* [S][TILE_BYTES] */
	IntraT* s_texc = reinterpret_cast<IntraT*>(ring + (size_t) S * TILE_BYTES);        /* [S][THREADS] lanes below, per slot */
	__shared__ __align__(8) u64 s_full[S];
	__shared__ u32 s_tile[NT];
	__shared__ AccT s_wsum[2][WARPS];                             /* by iteration parity: read after the barrier */
	__shared__ AccT s_wexc[S][WARPS];                             /* sums of the warps below, per ring slot */

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	/* this thread's 64 bytes: chunks 4 * tid + k of the tile = row tid / 2, columns 4 * (tid & 1) + k,
	 * stored at column ^ (row & 7) */
	const u32 row = (u32) tid >> 1;
	const u32 my_off = row * 128u;
	const u32 colx = (((u32) tid & 1u) << 2) ^ (row & 7u);
	auto chunk = [&](unsigned char* tile, int k) { return reinterpret_cast<uint4*>(tile + my_off + ((colx ^ (u32) k) << 4)); };

	u32 tk = 0;
	if (tid == 0) {
		for (int s = 0; s < S; ++s) stm_mbar_init(&s_full[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		stm_fence_async();
		const u32 t0 = atomicAdd(ticket, 1u);
		s_tile[0] = t0;
		tk = atomicAdd(ticket, 1u);
		if (t0 < num_tiles) {
			stm_mbar_expect_tx(&s_full[0], TILE_BYTES);
			stm_tma_load(ring, &tm_in, 0, (int) (t0 * (u32) ROWS), &s_full[0]);
		}
	}
	__syncthreads();
	unsigned waited = 0;
	for (u32 it = 0;; ++it) {
		const int r = (int) it - AHEAD_;                    /* sequence number reduced now */
		const int q = r - LAG_;                              /* sequence number scanned now */
		/* written before the barrier of the previous iteration */
		const u32 t_red = r >= 0 ? s_tile[r % NT] : 0xffffffffu;
		const u32 t_scan = q >= 0 ? s_tile[q % NT] : 0xffffffffu;
		if (q >= 0 && t_scan >= num_tiles) break;            /* tickets are monotonic: nothing left */
		/* the tile's prefix word: requested now, needed after the reduce below */
		AccT pfx = AccT(0);
		bool pfx_ok = false;
		if (lane == 0 && q >= 0) pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);

		/* (1) reduce the tile that has landed: thread sum, warp scan of the thread sums */
		const bool red = r >= 0 && t_red < num_tiles;
		if (red) {
			const int slot = r % S;
			if (!stm_mbar_wait(&s_full[slot], (u32) (r / S) & 1u) && lane == 0) atomicExch(err_flag, 1);
			unsigned char* tile = ring + (size_t) slot * TILE_BYTES;
			IntraT part[4];
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				ElemT e[EPV];
				*reinterpret_cast<uint4*>(e) = *chunk(tile, k);
				IntraT s2 = to_acc<ElemT, SumT, IntraT>(e[0]);
#pragma unroll
				for (int c = 1; c < EPV; ++c) s2 += to_acc<ElemT, SumT, IntraT>(e[c]);
				part[k] = s2;
			}
			const IntraT sum = (part[0] + part[1]) + (part[2] + part[3]);
			const IntraT incl = warp_inclusive_scan<IntraT>(sum, lane);
			IntraT excl = __shfl_up_sync(0xffffffffu, incl, 1);
			if (lane == 0) excl = IntraT(0);
			s_texc[slot * THREADS + tid] = excl;
			if (lane == 31) s_wsum[it & 1][warp] = static_cast<AccT>(incl);
		}

		/* (2) scan the tile whose prefix was requested above, in place */
		if (q >= 0) {
			const int slot = q % S;
			if (lane == 0) {
				unsigned spins = 0;
				while (!pfx_ok) {
					if (++spins > (1u << 22)) { atomicExch(err_flag, 1); break; }
					if (!(flags & 4)) __nanosleep(100);              /* do not stand in the propagator's way */
					pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);
				}
				waited += spins;
			}
			const AccT b = (__shfl_sync(0xffffffffu, pfx, 0) + s_wexc[slot][warp]) + static_cast<AccT>(s_texc[slot * THREADS + tid]);
			unsigned char* tile = ring + (size_t) slot * TILE_BYTES;
			if (FLOAT32) {
				/* the f64 base is rounded once; the thread-local exclusive sums stay small */
				const IntraT bf = static_cast<IntraT>(b);
				IntraT run = IntraT(0);
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					uint4* p = chunk(tile, k);
					ElemT e[EPV];
					SumT o[EPV];
					*reinterpret_cast<uint4*>(e) = *p;
#pragma unroll
					for (int c = 0; c < EPV; ++c) {
						o[c] = static_cast<SumT>(bf + run);
						run += to_acc<ElemT, SumT, IntraT>(e[c]);
					}
					*p = *reinterpret_cast<uint4*>(o);
				}
			} else {
				AccT run = b;
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					uint4* p = chunk(tile, k);
					ElemT e[EPV];
					SumT o[EPV];
					*reinterpret_cast<uint4*>(e) = *p;
#pragma unroll
					for (int c = 0; c < EPV; ++c) {
						o[c] = static_cast<SumT>(run);
						run += to_acc<ElemT, SumT, AccT>(e[c]);
					}
					*p = *reinterpret_cast<uint4*>(o);
				}
			}
			stm_fence_async();                               /* my writes -> visible to the copy engine */
		}
		__syncthreads();                                     /* the only barrier of the iteration */
		if (tid == 0) {
			if (q >= 0) {
				stm_tma_store(&tm_out, ring + (size_t) (q % S) * TILE_BYTES, 0, (int) (t_scan * (u32) ROWS));
				stm_bulk_commit();
			}
			/* next tile: ticket drawn one iteration ago; its slot was stored from S - AHEAD - LAG - 1
			 * iterations ago at the latest */
			const u32 t_next = tk;
			s_tile[(it + 1) % NT] = t_next;
			tk = atomicAdd(ticket, 1u);
			if (t_next < num_tiles) {
				const int slot = (int) ((it + 1) % S);
				stm_bulk_wait_read<SLACK>();
				stm_mbar_expect_tx(&s_full[slot], TILE_BYTES);
				stm_tma_load(ring + (size_t) slot * TILE_BYTES, &tm_in, 0, (int) (t_next * (u32) ROWS), &s_full[slot]);
			}
		} else if (tid == 32 && red) {
			AccT total = AccT(0);
#pragma unroll
			for (int w = 0; w < WARPS; ++w) { s_wexc[r % S][w] = total; total += s_wsum[it & 1][w]; }
			spp_publish<AccT>(agg + (size_t) t_red * AW::N, epoch, total);
		}
	}
	if (tid == 0) stm_bulk_wait_read<0>();
	if (waited) atomicAdd(reinterpret_cast<unsigned*>(err_flag) + 2, waited);   /* statistics: prefix polls */
}

#endif
