/*
 * scan_pp.cuh -- the single-pass scan as a PERSISTENT kernel with a shared-memory tile
 * ring and a PREFIX-PROPAGATOR CTA.  Included by scan.cu (uses its traits and helpers).
 *
 * The one-tile-per-CTA kernel (clo_scan_lookback) keeps 64 warps per SM resident but ncu
 * shows 34 of 52 stall cycles per issue at the barrier: every CTA idles, holding its
 * registers, while one warp walks back over unresolved predecessors (a loaded L2 round trip
 * is ~2000 cycles and a tile is issued chip-wide every few cycles), and nobody is loading.
 * Here a worker CTA
 *   - streams tiles into a ring of STAGES shared-memory slots with cp.async (each thread
 *     only ever touches the bytes it copied itself, so the ring needs no barriers),
 *   - reduces a tile as soon as it has landed and publishes the tile total (AGG),
 *   - scans and stores a tile LAG iterations later, when its exclusive prefix (PREF) has
 *     long been published by the propagator CTA, which consumes AGG words strictly in tile
 *     order and publishes PREF[t] as soon as AGG[0..t-1] are there (no dependency on later
 *     tiles, hence no deadlock; tiles are handed out by an atomic ticket).
 * AGG/PREF words carry the call's epoch, so nothing is reset between calls.
 * HBM traffic is unchanged: every element is read once and written once.
 */
#ifndef CLO_SCAN_PP_CUH
#define CLO_SCAN_PP_CUH

const int SPP_THREADS = 256;
const int SPP_VPT = 4;          /* 16-byte vectors per thread per tile */
const int SPP_LAG = 4;          /* iterations between a tile's reduce and its scan (hides the propagation) */
const int SPP_AHEAD = 1;        /* tiles in flight (cp.async) ahead of the reduce */
const int SPP_STAGES = SPP_AHEAD + 1 + SPP_LAG;
const int SPP_PU = 4;           /* propagator: tiles per thread per iteration */

__device__ __forceinline__ void spp_cp_async16(void* smem_dst, const void* gmem_src) {
	const unsigned s = (unsigned) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void spp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void spp_cp_async_wait() {
	asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

/* epoch-tagged descriptor: AW::N words of (epoch << 32 | payload32) */
template <typename AccT>
__device__ __forceinline__ void spp_publish(u64* p, u32 epoch, AccT v) {
	u32 w[AccWords<AccT>::N];
	AccWords<AccT>::pack(v, w);
#pragma unroll
	for (int k = 0; k < AccWords<AccT>::N; ++k) st_relaxed(p + k, ((u64) epoch << 32) | w[k]);
}
template <typename AccT>
__device__ __forceinline__ bool spp_read(const u64* p, u32 epoch, AccT& v) {
	u32 w[AccWords<AccT>::N];
	bool ok = true;
#pragma unroll
	for (int k = 0; k < AccWords<AccT>::N; ++k) {
		const u64 x = ld_relaxed(p + k);
		w[k] = (u32) x;
		if ((u32) (x >> 32) != epoch) ok = false;
	}
	v = AccWords<AccT>::unpack(w);
	return ok;
}

/* ---- propagator: PREF[t] = carry + sum of AGG[0..t-1], in tile order, streaming */
template <typename AccT, int THREADS>
__device__ __forceinline__ void spp_propagate(u64* __restrict__ agg, u64* __restrict__ pref, u32 num_tiles,
		u32 epoch, AccT carry, int* __restrict__ err_flag) {
	typedef AccWords<AccT> AW;
	constexpr int WARPS = THREADS / 32;
	constexpr int R = THREADS * SPP_PU;        /* window */
	__shared__ AccT s_w[WARPS];
	__shared__ int s_first[2];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	AccT running = carry;
	if (tid < 2) s_first[tid] = R;
	if (tid == 0) spp_publish<AccT>(pref, epoch, carry);      /* PREF[0] */
	__syncthreads();
	u32 t0 = 0;
	int par = 0;
	unsigned idle = 0;
	AccT v[SPP_PU];
	bool ok[SPP_PU];
	auto load_window = [&](u32 base) {
#pragma unroll
		for (int u = 0; u < SPP_PU; ++u) {
			const u32 t = base + (u32) tid * SPP_PU + u;      /* thread-contiguous tiles */
			v[u] = AccT(0);
			ok[u] = (t < num_tiles) && spp_read<AccT>(agg + (size_t) t * AW::N, epoch, v[u]);
		}
	};
	load_window(0);
	while (t0 < num_tiles) {
		int first_bad = R;
#pragma unroll
		for (int u = SPP_PU - 1; u >= 0; --u) if (!ok[u]) first_bad = tid * SPP_PU + u;
#pragma unroll
		for (int off = 16; off > 0; off >>= 1) first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, off));
		if (lane == 0 && first_bad < R) atomicMin(&s_first[par], first_bad);
		/* thread-local inclusive sums, warp scan of the thread totals */
		AccT inc[SPP_PU];
		inc[0] = v[0];
#pragma unroll
		for (int u = 1; u < SPP_PU; ++u) inc[u] = inc[u - 1] + v[u];
		const AccT wincl = warp_inclusive_scan<AccT>(inc[SPP_PU - 1], lane);
		if (lane == 31) s_w[warp] = wincl;
		__syncthreads();
		const int L = s_first[par];            /* tiles t0 .. t0+L-1 are there */
		if (tid == 0) s_first[par ^ 1] = R;
		/* exclusive offset of this thread inside the window */
		AccT toff = wincl - inc[SPP_PU - 1];
#pragma unroll
		for (int w = 0; w < WARPS; ++w) if (w < warp) toff += s_w[w];
		/* total of the consumed part (entries < L): held by the thread that owns entry L-1 */
		/* next window's loads go out before this iteration's stores */
		AccT pv[SPP_PU];
#pragma unroll
		for (int u = 0; u < SPP_PU; ++u) pv[u] = toff + inc[u];
		const u32 base = t0;
		load_window(t0 + (u32) L);
		if (L == 0) {
			if (++idle > (1u << 22)) { atomicExch(err_flag, 1); break; }
			par ^= 1;
			__syncthreads();
			continue;
		}
#pragma unroll
		for (int u = 0; u < SPP_PU; ++u) {
			const int j = tid * SPP_PU + u;
			if (j < L && base + j + 1 < num_tiles)
				spp_publish<AccT>(pref + (size_t) (base + j + 1) * AW::N, epoch, running + pv[u]);
		}
		/* everyone needs the consumed total: broadcast through shared memory */
		__shared__ AccT s_total;
		{
			const int jl = L - 1;
			if (tid == jl / SPP_PU) {
				AccT t = AccT(0);
#pragma unroll
				for (int u = 0; u < SPP_PU; ++u) if (u == jl % SPP_PU) t = pv[u];
				s_total = t;
			}
		}
		__syncthreads();
		running += s_total;
		t0 += (u32) L;
		par ^= 1;
		idle = 0;
		__syncthreads();
	}
}

/* ---- propagator, second form: a CHAIN OF WARPS.
 * The windowed propagator above pays, per window, an L2 round trip for the loads it can only
 * issue after the previous window's length is known, three CTA barriers and a shared-memory
 * scan; an AGG word waits up to a whole window period before anybody looks at it.  Measured with
 * the copy-engine workers (scan_tma.cuh), whose tiles come 1.5 us apart: hundreds of prefix polls
 * per tile, and the polls themselves slow the propagator down further.
 * Here warp w owns the chunks w, w + WARPS, ... of 32 * PU consecutive tiles.  Its lanes poll
 * their own AGG words back to back (no barrier, no other warp involved), one warp scan turns
 * them into exclusive sums, and the chunk's base comes from the previous chunk's warp through a
 * shared-memory mailbox -- the only serial link, a few dozen cycles instead of an L2 round
 * trip.  A tile's PREF leaves about one poll period after the last AGG of its chunk arrived. */
template <typename AccT, int THREADS, int PU>
__device__ __forceinline__ void spp_propagate_chain(u64* __restrict__ agg, u64* __restrict__ pref, u32 num_tiles,
		u32 epoch, AccT carry, int* __restrict__ err_flag) {
	typedef AccWords<AccT> AW;
	constexpr int WARPS = THREADS / 32;
	constexpr u32 CH = 32 * PU;
	__shared__ AccT s_box[WARPS];
	__shared__ u32 s_seq[WARPS];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	if (tid < WARPS) s_seq[tid] = 0;
	__syncthreads();
	if (tid == 0) { s_box[0] = carry; __threadfence_block(); *reinterpret_cast<volatile u32*>(&s_seq[0]) = 1u; }
	const u32 num_chunks = (num_tiles + CH - 1) / CH;
	for (u32 c = (u32) warp; c < num_chunks; c += WARPS) {
		const u32 t0 = c * CH + (u32) lane * PU;
		AccT v[PU];
		bool ok[PU];
#pragma unroll
		for (int u = 0; u < PU; ++u) { v[u] = AccT(0); ok[u] = t0 + u >= num_tiles; }
		unsigned spins = 0;
		for (;;) {
#pragma unroll
			for (int u = 0; u < PU; ++u)
				if (!ok[u]) ok[u] = spp_read<AccT>(agg + (size_t) (t0 + u) * AW::N, epoch, v[u]);
			bool all = true;
#pragma unroll
			for (int u = 0; u < PU; ++u) all = all && ok[u];
			if (__all_sync(0xffffffffu, all)) break;
			if (++spins > (1u << 22)) { if (lane == 0) atomicExch(err_flag, 1); break; }
		}
		AccT inc[PU];
		inc[0] = v[0];
#pragma unroll
		for (int u = 1; u < PU; ++u) inc[u] = inc[u - 1] + v[u];
		const AccT wincl = warp_inclusive_scan<AccT>(inc[PU - 1], lane);
		const AccT total = __shfl_sync(0xffffffffu, wincl, 31);
		AccT base = AccT(0);
		if (lane == 0) {
			unsigned w2 = 0;
			while (*reinterpret_cast<volatile u32*>(&s_seq[warp]) != c + 1u) {
				if (++w2 > (1u << 26)) { atomicExch(err_flag, 1); break; }
			}
			__threadfence_block();
			base = *reinterpret_cast<volatile AccT*>(&s_box[warp]);
			const int nw = warp + 1 == WARPS ? 0 : warp + 1;
			*reinterpret_cast<volatile AccT*>(&s_box[nw]) = base + total;
			__threadfence_block();
			*reinterpret_cast<volatile u32*>(&s_seq[nw]) = c + 2u;
		}
		base = __shfl_sync(0xffffffffu, base, 0);
		const AccT lane_excl = base + (wincl - inc[PU - 1]);
#pragma unroll
		for (int u = 0; u < PU; ++u)
			if (t0 + u < num_tiles)
				spp_publish<AccT>(pref + (size_t) (t0 + u) * AW::N, epoch, u == 0 ? lane_excl : lane_excl + inc[u - 1]);
	}
}

/* ---- the kernel: block 0 propagates, the others work */
template <typename ElemT, typename SumT, int THREADS, int VPT_ = SPP_VPT, int AHEAD_ = SPP_AHEAD, int LAG_ = SPP_LAG>
__global__ void __launch_bounds__(THREADS, 2)
clo_scan_pp(const ElemT* __restrict__ in, SumT* __restrict__ out, size_t n, u32 num_tiles,
		u64* __restrict__ agg, u64* __restrict__ pref, u32* __restrict__ ticket, u32 epoch,
		const SumT* __restrict__ carry_in, int* __restrict__ err_flag) {
	typedef typename AccOf<SumT>::type AccT;
	typedef AccWords<AccT> AW;
	/* tile-local arithmetic: f32 for f32 sums (a 4096-element shuffle/tree sum loses ~1e-6
	 * relative, far inside the stated tolerance); everything between tiles stays in AccT (f64) */
	typedef typename std::conditional<std::is_same<SumT, float>::value, float, AccT>::type IntraT;
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;          /* elements per vector */
	constexpr int VB = EPV * (int) sizeof(ElemT);             /* vector bytes: 4, 8 or 16 */
	static_assert(VB == 16, "the ring is filled with 16-byte cp.async");
	constexpr int VPT = VPT_;
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * VPT * EPV;
	constexpr int S = AHEAD_ + 1 + LAG_;
	constexpr int OCH = (sizeof(SumT) * EPV <= 16) ? EPV : (16 / (int) sizeof(SumT));

	const AccT carry = carry_in ? to_acc<SumT, SumT, AccT>(*carry_in) : AccT(0);
	if (blockIdx.x == 0) {
		spp_propagate_chain<AccT, THREADS, 1>(agg, pref, num_tiles, epoch, carry, err_flag);
		return;
	}

	extern __shared__ __align__(16) unsigned char spp_smem[];
	ElemT* ring = reinterpret_cast<ElemT*>(spp_smem);            /* [S][TILE] */
	__shared__ u32 s_tile[S];
	__shared__ AccT s_wsum[S][WARPS];
	__shared__ AccT s_pref;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	/* element offset of vector j of this thread inside a tile (warp-striped rows) */
	const u32 lane_off = ((u32) (warp * VPT) * 32u + lane) * EPV;

	/* the ticket is always drawn one iteration early so that its round trip is hidden */
	u32 tk = 0;
	if (tid == 0) tk = atomicAdd(ticket, 1u);
	for (u32 it = 0;; ++it) {
		const int slot_load = (int) (it % S);
		/* (1) ticket for the tile loaded this iteration */
		if (tid == 0) s_tile[slot_load] = tk;
		const int r = (int) it - AHEAD_;                    /* sequence number reduced now */
		const int q = r - LAG_;                              /* sequence number scanned now */
		const u32 t_scan = q >= 0 ? s_tile[q % S] : 0xffffffffu;   /* read before it can be overwritten */
		__syncthreads();
		const u32 t_load = s_tile[slot_load];
		if (q >= 0 && t_scan >= num_tiles) break;            /* tickets are monotonic: nothing left */
		if (tid == 0) tk = atomicAdd(ticket, 1u);
		/* the prefix word of the tile scanned below is requested now, used after two barriers */
		AccT pfx = AccT(0);
		bool pfx_ok = false;
		if (tid == 0 && q >= 0) pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);

		/* (2) start the copy of the new tile (own bytes only); always commit a group */
		if (t_load < num_tiles) {
			const size_t base = (size_t) t_load * TILE;
			ElemT* dst = ring + (size_t) slot_load * TILE;
			if (base + TILE <= n) {
#pragma unroll
				for (int j = 0; j < VPT; ++j) {
					const u32 o = lane_off + (u32) j * 32 * EPV;
					spp_cp_async16(dst + o, in + base + o);
				}
			} else {
#pragma unroll
				for (int j = 0; j < VPT; ++j)
#pragma unroll
					for (int c = 0; c < EPV; ++c) {
						const u32 o = lane_off + (u32) j * 32 * EPV + c;
						dst[o] = (base + o < n) ? in[base + o] : ElemT(0);
					}
			}
		}
		spp_cp_async_commit();

		/* (3) reduce the tile that has landed, publish its total */
		const u32 t_red = r >= 0 ? s_tile[r % S] : 0xffffffffu;
		if (r >= 0 && t_red < num_tiles) {
			spp_cp_async_wait<AHEAD_>();
			const ElemT* src = ring + (size_t) (r % S) * TILE;
			IntraT sum = IntraT(0);
#pragma unroll
			for (int j = 0; j < VPT; ++j) {
				ElemT e[EPV];
				*reinterpret_cast<uint4*>(e) = *reinterpret_cast<const uint4*>(src + lane_off + (u32) j * 32 * EPV);
#pragma unroll
				for (int c = 0; c < EPV; ++c) sum += to_acc<ElemT, SumT, IntraT>(e[c]);
			}
			sum = warp_reduce_sum<IntraT>(sum);
			if (lane == 0) s_wsum[r % S][warp] = static_cast<AccT>(sum);
		}
		__syncthreads();
		if (r >= 0 && t_red < num_tiles && tid == 0) {
			AccT total = AccT(0);
#pragma unroll
			for (int w = 0; w < WARPS; ++w) total += s_wsum[r % S][w];
			spp_publish<AccT>(agg + (size_t) t_red * AW::N, epoch, total);
		}

		/* (4) scan + store the tile whose prefix was requested LAG iterations ago */
		if (q >= 0) {
			if (tid == 0) {
				unsigned spins = 0;
				while (!pfx_ok) {
					if (++spins > (1u << 24)) { atomicExch(err_flag, 1); break; }
					pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);
				}
				s_pref = pfx;
			}
			__syncthreads();
			const ElemT* src = ring + (size_t) (q % S) * TILE;
			const size_t base = (size_t) t_scan * TILE;
			const bool full = base + TILE <= n;
			AccT warp_off = s_pref;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) if (w < warp) warp_off += s_wsum[q % S][w];
			AccT row_off = warp_off;
#pragma unroll
			for (int j = 0; j < VPT; ++j) {
				const u32 o = lane_off + (u32) j * 32 * EPV;
				ElemT e[EPV];
				*reinterpret_cast<uint4*>(e) = *reinterpret_cast<const uint4*>(src + o);
				IntraT v[EPV];
#pragma unroll
				for (int c = 0; c < EPV; ++c) v[c] = to_acc<ElemT, SumT, IntraT>(e[c]);
#pragma unroll
				for (int c = 1; c < EPV; ++c) v[c] += v[c - 1];
				const IntraT incl = warp_inclusive_scan<IntraT>(v[EPV - 1], lane);
				IntraT excl = __shfl_up_sync(0xffffffffu, incl, 1);
				if (lane == 0) excl = IntraT(0);
				const IntraT row_total = __shfl_sync(0xffffffffu, incl, 31);
				const AccT b = row_off + static_cast<AccT>(excl);
				row_off += static_cast<AccT>(row_total);
				SumT ov[EPV];
				if (std::is_same<SumT, float>::value) {
					/* round the f64 base once per vector; one more f32 rounding per element */
					const IntraT bf = static_cast<IntraT>(b);
					ov[0] = static_cast<SumT>(bf);
#pragma unroll
					for (int c = 1; c < EPV; ++c) ov[c] = static_cast<SumT>(bf + v[c - 1]);
				} else {
					ov[0] = static_cast<SumT>(b);
#pragma unroll
					for (int c = 1; c < EPV; ++c) ov[c] = static_cast<SumT>(b + static_cast<AccT>(v[c - 1]));
				}
				const size_t idx = base + o;
				if (full || idx + EPV <= n) {
#pragma unroll
					for (int c0 = 0; c0 < EPV; c0 += OCH) {
						SumT chunk[OCH];
#pragma unroll
						for (int c = 0; c < OCH; ++c) chunk[c] = ov[c0 + c];
						store_vec_cs<SumT, OCH>(out + idx + c0, chunk);
					}
				} else {
#pragma unroll
					for (int c = 0; c < EPV; ++c) if (idx + c < n) out[idx + c] = ov[c];
				}
			}
		}
	}
	spp_cp_async_wait<0>();
}

/* ---- second schedule of the same pipeline: ONE CTA barrier per tile.
 * ncu on the float instance of clo_scan_pp (profiles/r02_scan_pp_f32_ncu_summary.txt): 25 %
 * occupancy, issue slots 32 % busy, 3.9 of 11 stall cycles per issue at the CTA barrier and 3.7
 * on the short scoreboard -- a latency chain, not a bandwidth limit: three barriers per tile, the
 * third one behind thread 0 polling the tile's prefix word.  Here
 *   - the ticket of a tile is written to shared memory one iteration before it is read, and a
 *     tile's warp sums are read one barrier after they were written, so one barrier per iteration
 *     orders everything;
 *   - every warp reads the tile's PREF word itself (lane 0 requests it before the reduce, polls
 *     after it, broadcasts by shuffle): no warp waits for another warp's poll. */
template <typename ElemT, typename SumT, int THREADS, int VPT_ = SPP_VPT, int AHEAD_ = SPP_AHEAD, int LAG_ = SPP_LAG>
__global__ void __launch_bounds__(THREADS, 2)
clo_scan_pp1b(const ElemT* __restrict__ in, SumT* __restrict__ out, size_t n, u32 num_tiles,
		u64* __restrict__ agg, u64* __restrict__ pref, u32* __restrict__ ticket, u32 epoch,
		const SumT* __restrict__ carry_in, int* __restrict__ err_flag) {
	typedef typename AccOf<SumT>::type AccT;
	typedef AccWords<AccT> AW;
	typedef typename std::conditional<std::is_same<SumT, float>::value, float, AccT>::type IntraT;
	constexpr int EPV = sizeof(ElemT) >= 8 ? 2 : 4;
	constexpr int VB = EPV * (int) sizeof(ElemT);
	static_assert(VB == 16, "the ring is filled with 16-byte cp.async");
	static_assert(AHEAD_ >= 1, "a tile's warp sums must not be rewritten in the iteration after their last use");
	constexpr int VPT = VPT_;
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * VPT * EPV;
	constexpr int S = AHEAD_ + 1 + LAG_;
	constexpr int NT = S + 2;                                 /* ticket slots */
	constexpr int OCH = (sizeof(SumT) * EPV <= 16) ? EPV : (16 / (int) sizeof(SumT));

	const AccT carry = carry_in ? to_acc<SumT, SumT, AccT>(*carry_in) : AccT(0);
	if (blockIdx.x == 0) {
		spp_propagate_chain<AccT, THREADS, 1>(agg, pref, num_tiles, epoch, carry, err_flag);
		return;
	}

	extern __shared__ __align__(16) unsigned char spp_smem[];
	ElemT* ring = reinterpret_cast<ElemT*>(spp_smem);            /* [S][TILE] */
	__shared__ u32 s_tile[NT];
	__shared__ AccT s_wsum[2][WARPS];                             /* by iteration parity: read by thread 0 after the barrier */
	__shared__ AccT s_wexc[S][WARPS];                             /* sums of the warps below, per ring slot */

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const u32 lane_off = ((u32) (warp * VPT) * 32u + lane) * EPV;

	/* prologue: ticket of sequence 0 in shared memory, ticket of sequence 1 in flight */
	u32 tk = 0;
	if (tid == 0) { s_tile[0] = atomicAdd(ticket, 1u); tk = atomicAdd(ticket, 1u); }
	__syncthreads();
	unsigned waited = 0;
	for (u32 it = 0;; ++it) {
		const int r = (int) it - AHEAD_;                    /* sequence number reduced now */
		const int q = r - LAG_;                              /* sequence number scanned now */
		/* all of these were written before the barrier of the previous iteration */
		const u32 t_load = s_tile[it % NT];
		const u32 t_red = r >= 0 ? s_tile[r % NT] : 0xffffffffu;
		const u32 t_scan = q >= 0 ? s_tile[q % NT] : 0xffffffffu;
		if (q >= 0 && t_scan >= num_tiles) break;            /* tickets are monotonic: nothing left */
		if (tid == 0) { s_tile[(it + 1) % NT] = tk; tk = atomicAdd(ticket, 1u); }
		/* the prefix word of the tile scanned below: requested now, looked at after the barrier */
		AccT pfx = AccT(0);
		bool pfx_ok = false;
		if (lane == 0 && q >= 0) pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);

		/* (1) start the copy of the new tile (own bytes only); always commit a group */
		if (t_load < num_tiles) {
			const size_t base = (size_t) t_load * TILE;
			ElemT* dst = ring + (size_t) (it % S) * TILE;
			if (base + TILE <= n) {
#pragma unroll
				for (int j = 0; j < VPT; ++j) {
					const u32 o = lane_off + (u32) j * 32 * EPV;
					spp_cp_async16(dst + o, in + base + o);
				}
			} else {
#pragma unroll
				for (int j = 0; j < VPT; ++j)
#pragma unroll
					for (int c = 0; c < EPV; ++c) {
						const u32 o = lane_off + (u32) j * 32 * EPV + c;
						dst[o] = (base + o < n) ? in[base + o] : ElemT(0);
					}
			}
		}
		spp_cp_async_commit();

		/* (2) reduce the tile that has landed */
		const bool red = r >= 0 && t_red < num_tiles;
		if (red) {
			spp_cp_async_wait<AHEAD_>();
			const ElemT* src = ring + (size_t) (r % S) * TILE;
			IntraT part[VPT];
#pragma unroll
			for (int j = 0; j < VPT; ++j) {
				ElemT e[EPV];
				*reinterpret_cast<uint4*>(e) = *reinterpret_cast<const uint4*>(src + lane_off + (u32) j * 32 * EPV);
				IntraT s2 = IntraT(0);
#pragma unroll
				for (int c = 0; c < EPV; ++c) s2 += to_acc<ElemT, SumT, IntraT>(e[c]);
				part[j] = s2;
			}
			IntraT sum = part[0];
#pragma unroll
			for (int j = 1; j < VPT; ++j) sum += part[j];
			sum = warp_reduce_sum<IntraT>(sum);
			if (lane == 0) s_wsum[it & 1][warp] = static_cast<AccT>(sum);
		}
		__syncthreads();                                     /* the only barrier of the iteration */
		if (red && tid == 0) {
			AccT total = AccT(0);
#pragma unroll
			for (int w = 0; w < WARPS; ++w) { s_wexc[r % S][w] = total; total += s_wsum[it & 1][w]; }
			spp_publish<AccT>(agg + (size_t) t_red * AW::N, epoch, total);
		}

		/* (3) scan + store the tile whose prefix was requested above */
		if (q >= 0) {
			if (lane == 0) {
				unsigned spins = 0;
				while (!pfx_ok) {
					if (++spins > (1u << 24)) { atomicExch(err_flag, 1); break; }
					pfx_ok = spp_read<AccT>(pref + (size_t) t_scan * AW::N, epoch, pfx);
				}
				waited += spins;
			}
			/* lane 0 holds the tile's prefix; the sums of the warps below were prefix-summed by
			 * thread 0 when the tile was reduced */
			AccT warp_off = __shfl_sync(0xffffffffu, pfx, 0) + s_wexc[q % S][warp];
			const ElemT* src = ring + (size_t) (q % S) * TILE;
			const size_t base = (size_t) t_scan * TILE;
			const bool full = base + TILE <= n;
			AccT row_off = warp_off;
#pragma unroll
			for (int j = 0; j < VPT; ++j) {
				const u32 o = lane_off + (u32) j * 32 * EPV;
				ElemT e[EPV];
				*reinterpret_cast<uint4*>(e) = *reinterpret_cast<const uint4*>(src + o);
				IntraT v[EPV];
#pragma unroll
				for (int c = 0; c < EPV; ++c) v[c] = to_acc<ElemT, SumT, IntraT>(e[c]);
#pragma unroll
				for (int c = 1; c < EPV; ++c) v[c] += v[c - 1];
				const IntraT incl = warp_inclusive_scan<IntraT>(v[EPV - 1], lane);
				IntraT excl = __shfl_up_sync(0xffffffffu, incl, 1);
				if (lane == 0) excl = IntraT(0);
				const IntraT row_total = __shfl_sync(0xffffffffu, incl, 31);
				const AccT b = row_off + static_cast<AccT>(excl);
				row_off += static_cast<AccT>(row_total);
				SumT ov[EPV];
				if (std::is_same<SumT, float>::value) {
					const IntraT bf = static_cast<IntraT>(b);
					ov[0] = static_cast<SumT>(bf);
#pragma unroll
					for (int c = 1; c < EPV; ++c) ov[c] = static_cast<SumT>(bf + v[c - 1]);
				} else {
					ov[0] = static_cast<SumT>(b);
#pragma unroll
					for (int c = 1; c < EPV; ++c) ov[c] = static_cast<SumT>(b + static_cast<AccT>(v[c - 1]));
				}
				const size_t idx = base + o;
				if (full || idx + EPV <= n) {
#pragma unroll
					for (int c0 = 0; c0 < EPV; c0 += OCH) {
						SumT chunk[OCH];
#pragma unroll
						for (int c = 0; c < OCH; ++c) chunk[c] = ov[c0 + c];
						store_vec_cs<SumT, OCH>(out + idx + c0, chunk);
					}
				} else {
#pragma unroll
					for (int c = 0; c < EPV; ++c) if (idx + c < n) out[idx + c] = ov[c];
				}
			}
		}
	}
	spp_cp_async_wait<0>();
	if (waited) atomicAdd(reinterpret_cast<unsigned*>(err_flag) + 2, waited);   /* statistics: prefix polls */
}

#endif
