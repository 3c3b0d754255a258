/*
 * radix_v6.cuh -- the onesweep pass for keys-only sorts, restructured around what the
 * B200 SM actually charges for (tools/ubench_smem.cu: a warp instruction on random digits
 * costs 2.4-2.8 SM cycles in the shared-memory pipe, a linear one 1.0; the old kernel spent
 * 44 % of a tile's time in phases that issue almost none of them, behind 9 CTA barriers).
 * Included by radix.cu; shares the prefix-propagator CTAs and the AGG/PREF protocol of
 * radix_pp.cuh.
 *
 * Per tile a worker CTA now runs two barriers instead of nine:
 *   P1 count   every key does one RED on its warp's private histogram (no result, so no
 *              registers are held for ranks)
 *   -- B1 --
 *   P2 digits  threads 0..255   column sums -> publish AGG[tile] -> tile-local digit starts
 *                               -> the histogram rows are turned into per-warp BASES
 *              threads 256..511 meanwhile wait for PREF of the PREVIOUS tile and build its
 *                               global offset table; one of them draws the next ticket
 *   -- B2 --
 *   P3 place   p = atomicAdd(&base[digit], 1) is the key's slot in the staged tile (stable
 *              across warps by construction, inside a warp instruction iff same-address
 *              lanes are served in lane order -- verified below, never trusted); every
 *              warp then clears its own histogram row (warp-private: no barrier)
 *   P4 load    keys of the next tile -> registers (in flight during P5 and the next P1 wait)
 *   P5 write   the PREVIOUS tile leaves its staging buffer, coalesced, and is verified
 *              while it does: (key & bits-sorted-so-far) must be non-decreasing along the
 *              staged order.  A failing tile is repaired after the next B1 with ballot ranks
 *              (slow, self-contained, never seen on B200; forced by flag 8 in the tests).
 * The staging buffers alternate, so a tile's write-out is separated from its staging by the
 * two barriers of the following iteration and needs none of its own.
 */
#ifndef CLO_RADIX_V6_CUH
#define CLO_RADIX_V6_CUH

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

template <typename ElemT>
__device__ __forceinline__ u32 v6_digit(ElemT k, u32 start_bit, u32 dmask) {
	if (sizeof(ElemT) <= 4) return (((u32) k) >> start_bit) & dmask;
	return ((u32) (((u64) k) >> start_bit)) & dmask;
}

template <typename ElemT, typename LbT, int THREADS, int IPT>
__global__ void __launch_bounds__(THREADS, 2)
clo_radix_onesweep_v6(const ElemT* __restrict__ in, ElemT* __restrict__ out, size_t n, u32 num_tiles,
		LbT* __restrict__ agg, LbT* __restrict__ pref, u32* __restrict__ ticket,
		const u64* __restrict__ bins_base, u32 start_bit, u32 dmask,
		int* __restrict__ err_flag, int prof_on, int flags) {
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * IPT;
	constexpr u32 NONE = 0xffffffffu;
	static_assert(THREADS >= 2 * RADIX, "256 digit threads + 256 prefix threads");
	static_assert((size_t) TILE * sizeof(ElemT) >= (size_t) WARPS * RADIX * 4, "repair scratch lives in a staging buffer");

	extern __shared__ __align__(16) unsigned char smem_raw[];

	if (blockIdx.x < PP_NUM_PROP) {
		pp_propagate<LbT, THREADS>(agg, pref, num_tiles, err_flag, smem_raw, prof_on);
		return;
	}

	u32* whist = reinterpret_cast<u32*>(smem_raw);                          /* [WARPS][RADIX] */
	u32* s_ds = whist + WARPS * RADIX;                                      /* [2][RADIX] digit starts */
	u64* s_goff_raw = reinterpret_cast<u64*>(s_ds + 2 * RADIX);             /* [2][RADIX] u64-sized slots */
	u32* s_misc = reinterpret_cast<u32*>(s_goff_raw + 2 * RADIX);           /* [16]: 0..7 scan, 8 ticket, 10..11 bad */
	ElemT* s_buf = reinterpret_cast<ElemT*>(s_misc + 16);                   /* [2][TILE] */

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const ElemT low_mask = (ElemT) ((((ElemT) dmask) << start_bit) | ((((ElemT) 1) << start_bit) - 1));
	const u32 wbase = (u32) warp * 32u * IPT + lane;
	u32* wh = whist + warp * RADIX;
	/* the prefix threads own one digit each */
	const LbT my_base = tid >= RADIX && tid < 2 * RADIX ? (LbT) bins_base[tid - RADIX] : (LbT) 0;

	ElemT key[IPT];

	/* optional phase profile: cycles per phase, thread 0 (digit side) and thread 256 (prefix side) */
	u64* prof = reinterpret_cast<u64*>(err_flag + 16);
	long long t_prev = 0;
	auto mark = [&](int phase) {
		if (prof_on && (tid == 0)) {
			const long long t = clock64();
			atomicAdd(prof + phase, (u64) (t - t_prev));
			t_prev = t;
		}
	};
	if (prof_on && tid == 0) t_prev = clock64();

	auto tile_count_of = [&](u32 t) -> u32 {
		const size_t base = (size_t) t * TILE;
		return (base + TILE <= n) ? (u32) TILE : (u32) (n - base);
	};
	auto goff_of = [&](int b) { return reinterpret_cast<LbT*>(s_goff_raw + (size_t) b * RADIX); };
	auto load_tile = [&](u32 t) {
		const size_t base = (size_t) t * TILE;
		const u32 cnt = tile_count_of(t);
		const ElemT* p = in + base + wbase;
		if (cnt == (u32) TILE) {
#pragma unroll
			for (int i = 0; i < IPT; ++i) key[i] = __ldcs(p + i * 32);
		} else {
#pragma unroll
			for (int i = 0; i < IPT; ++i) key[i] = (wbase + i * 32u < cnt) ? __ldcs(p + i * 32) : ElemT(0);
		}
	};
	auto zero_row = [&]() {
		uint4* row = reinterpret_cast<uint4*>(wh);
#pragma unroll
		for (int i = 0; i < RADIX / 4 / 32; ++i) row[lane + i * 32] = make_uint4(0u, 0u, 0u, 0u);
	};
	/* coalesced write-out of a staged tile; returns true when the staged order is not
	 * sorted on the bits processed so far (= some warp instruction's atomics were not
	 * applied in lane order) */
	auto write_out = [&](u32 t, int b, bool verify) -> bool {
		const u32 cnt = tile_count_of(t);
		const ElemT* skeys = s_buf + (size_t) b * TILE;
		const LbT* goff = goff_of(b);
		bool bad = false;
		if ((flags & 8) && (t % 5u) == 2u) {
			/* test hook: write this tile WRONG and report it, so that only a working repair
			 * path gives a sorted result */
			for (u32 j = tid; j < cnt; j += THREADS) {
				const ElemT k = skeys[j];
				out[goff[v6_digit<ElemT>(k, start_bit, dmask)] + (LbT) j] = (ElemT) ~k;
			}
			return true;
		}
		if (cnt == (u32) TILE) {
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 j = (u32) tid + i * THREADS;
				const ElemT k = skeys[j];
				const ElemT kp = skeys[j > 0 ? j - 1 : 0];
				if ((k & low_mask) < (kp & low_mask)) bad = true;
				out[goff[v6_digit<ElemT>(k, start_bit, dmask)] + (LbT) j] = k;
			}
		} else {
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 j = (u32) tid + i * THREADS;
				if (j < cnt) {
					const ElemT k = skeys[j];
					const ElemT kp = skeys[j > 0 ? j - 1 : 0];
					if ((k & low_mask) < (kp & low_mask)) bad = true;
					out[goff[v6_digit<ElemT>(k, start_bit, dmask)] + (LbT) j] = k;
				}
			}
		}
		return verify && bad;
	};
	/* prefix threads: PREF[t] -> global offset table of the tile staged in buffer b */
	auto prefix_to_goff = [&](u32 t, int b, LbT w) {
		const int d = tid - RADIX;
		LbT* p = pref + (size_t) t * RADIX + d;
		unsigned spins = 0;
		while (!(w & PPWord<LbT>::VALID)) {
			if (++spins > SPIN_LIMIT) { atomicExch(err_flag, 1); break; }
			w = ld_relaxed(p);
		}
		if (prof_on && d == 0) atomicAdd(prof + 8, (u64) spins);
		st_relaxed(p, (LbT) 0);                                  /* consumed: reset */
		goff_of(b)[d] = my_base + (w & PPWord<LbT>::VAL) - (LbT) s_ds[b * RADIX + d];
	};
	/* Rare path.  Tile t (staged in buffer b, already written out in a wrong intra-digit
	 * order) is ranked again with ballots and scattered straight to its final positions.
	 * Uses buffer b as scratch; goff and s_ds of buffer b are still those of tile t.
	 * Called by all threads between B1 and P2; clobbers key[]. */
	auto repair = [&](u32 t, int b) {
		u32* scratch = reinterpret_cast<u32*>(s_buf + (size_t) b * TILE);   /* [WARPS][RADIX] */
		const u32 cnt = tile_count_of(t);
		if (tid == 0) atomicAdd(err_flag + 1, 1);
		for (int i = tid; i < WARPS * RADIX; i += THREADS) scratch[i] = 0;
		__syncthreads();
		const size_t base = (size_t) t * TILE;
		u32* sw = scratch + warp * RADIX;
		u32 rank[IPT];
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			const bool valid = local < cnt;
			key[i] = valid ? in[base + local] : ElemT(0);
			const u32 d = v6_digit<ElemT>(key[i], start_bit, dmask);
			u32 peers = match_digit_ballot(d);
			peers &= __ballot_sync(0xffffffffu, valid);
			const u32 lt = peers & lanemask_lt();
			u32 old = 0;
			if (valid && lt == 0) { old = sw[d]; sw[d] = old + __popc(peers); }
			__syncwarp();
			old = __shfl_sync(0xffffffffu, old, (__ffs(peers) - 1) & 31);
			rank[i] = old + __popc(lt);
		}
		__syncthreads();
		if (tid < RADIX) {
			u32 run = 0;
			for (int w = 0; w < WARPS; ++w) { const u32 c = scratch[w * RADIX + tid]; scratch[w * RADIX + tid] = run; run += c; }
		}
		__syncthreads();
		const LbT* goff = goff_of(b);
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			if (local < cnt) {
				const u32 d = v6_digit<ElemT>(key[i], start_bit, dmask);
				out[goff[d] + (LbT) (s_ds[b * RADIX + d] + sw[d] + rank[i])] = key[i];
			}
		}
		__syncthreads();
	};

	/* ---- prologue */
	if (tid == 0) { s_misc[8] = atomicAdd(ticket, 1u); s_misc[10] = 0; s_misc[11] = 0; }
	zero_row();
	__syncthreads();
	u32 cur = s_misc[8];
	if (cur >= num_tiles) return;
	load_tile(cur);
	u32 prev = NONE, pprev = NONE;
	int b = 0;
	u32 it = 0;
	for (;;) {
		LbT wp = 0;
		const bool is_pref_thread = tid >= RADIX && tid < 2 * RADIX;
		if (prev != NONE && is_pref_thread) wp = ld_relaxed(pref + (size_t) prev * RADIX + (tid - RADIX));
		const u32 cnt = tile_count_of(cur);
		/* P1 count */
		if (cnt == (u32) TILE) {
#pragma unroll
			for (int i = 0; i < IPT; ++i) atomicAdd(&wh[v6_digit<ElemT>(key[i], start_bit, dmask)], 1u);
		} else {
#pragma unroll
			for (int i = 0; i < IPT; ++i)
				if (wbase + i * 32u < cnt) atomicAdd(&wh[v6_digit<ElemT>(key[i], start_bit, dmask)], 1u);
		}
		mark(0);
		__syncthreads();                                         /* B1 */
		mark(1);
		/* the write-out of the previous iteration (tile pprev, buffer b) reported a bad order */
		if (s_misc[10 + ((it + 1) & 1)]) {
			repair(pprev, b);
			load_tile(cur);
			if (tid == 0) s_misc[10 + ((it + 1) & 1)] = 0;
		}
		/* P2 */
		if (tid < RADIX) {
			u32 c[WARPS];
#pragma unroll
			for (int w = 0; w < WARPS; ++w) c[w] = whist[w * RADIX + tid];
			u32 count = 0;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) count += c[w];
			st_relaxed(agg + (size_t) cur * RADIX + tid, (LbT) (PPWord<LbT>::VALID | (LbT) count));
			const u32 incl = warp_inclusive_scan<u32>(count, lane);
			if (lane == 31) s_misc[warp] = incl;
			named_bar_sync(1, RADIX);
			u32 off = 0;
#pragma unroll
			for (int w = 0; w < RADIX / 32; ++w) if (w < warp) off += s_misc[w];
			u32 run = off + incl - count;
			s_ds[b * RADIX + tid] = run;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) { whist[w * RADIX + tid] = run; run += c[w]; }
		} else if (is_pref_thread) {
			u32 nt = 0;
			if (tid == RADIX) nt = atomicAdd(ticket, 1u);        /* in flight during the prefix wait */
			if (prev != NONE) prefix_to_goff(prev, b ^ 1, wp);
			if (tid == RADIX) s_misc[8] = nt;
		}
		mark(2);
		__syncthreads();                                         /* B2 */
		mark(3);
		/* P3 place */
		{
			ElemT* skeys = s_buf + (size_t) b * TILE;
			if (cnt == (u32) TILE) {
#pragma unroll
				for (int i = 0; i < IPT; ++i)
					skeys[atomicAdd(&wh[v6_digit<ElemT>(key[i], start_bit, dmask)], 1u)] = key[i];
			} else {
#pragma unroll
				for (int i = 0; i < IPT; ++i)
					if (wbase + i * 32u < cnt)
						skeys[atomicAdd(&wh[v6_digit<ElemT>(key[i], start_bit, dmask)], 1u)] = key[i];
			}
			__syncwarp();
			zero_row();
		}
		mark(4);
		/* P4 next tile */
		const u32 nxt = s_misc[8];
		const bool more = nxt < num_tiles;
		if (more) load_tile(nxt);
		mark(5);
		/* P5 write-out of the previous tile */
		if (prev != NONE) {
			if (write_out(prev, b ^ 1, true)) s_misc[10 + (it & 1)] = 1;
		}
		mark(6);
		pprev = prev;
		prev = cur;
		b ^= 1;
		++it;
		if (!more) break;
		cur = nxt;
	}
	/* ---- epilogue: `prev` is staged in buffer b^1; pprev was written in the last iteration */
	__syncthreads();
	if (s_misc[10 + ((it + 1) & 1)]) repair(pprev, b);
	if (tid >= RADIX && tid < 2 * RADIX) prefix_to_goff(prev, b ^ 1, (LbT) 0);
	__syncthreads();
	const bool bad = write_out(prev, b ^ 1, true);
	if (__syncthreads_or(bad ? 1 : 0)) repair(prev, b ^ 1);
}

template <typename ElemT, int THREADS, int IPT, typename LbT>
constexpr size_t onesweep_v6_smem() {
	constexpr size_t worker = (size_t) (THREADS / 32) * RADIX * 4 + 2 * RADIX * 4 + 2 * RADIX * 8 + 16 * 4 +
		2 * (size_t) THREADS * IPT * sizeof(ElemT);
	constexpr size_t prop = ((size_t) PP_WINDOW * 32 + 8 * 32) * sizeof(LbT) + 16;
	return worker > prop ? worker : prop;
}

#endif
