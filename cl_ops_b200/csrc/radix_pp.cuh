/*
 * radix_pp.cuh -- the onesweep pass as a PERSISTENT, SOFTWARE-PIPELINED kernel with
 * dedicated PREFIX-PROPAGATOR CTAs.  Included by radix.cu (uses its helpers).
 *
 * Why: in the classic decoupled look-back every tile walks back over its unresolved
 * predecessors.  On B200 a tile is issued chip-wide every ~30-50 cycles while one loaded
 * L2 round trip is ~2000 cycles, so tens of predecessors are always unresolved and the walk
 * (measured: ~5k of ~15k cycles per tile, CLO_RADIX_PROFILE=1) dominates; widening the
 * window only adds L2 traffic.  Here
 *   - a worker CTA publishes the 256 digit counts of its tile (AGG) and never looks back;
 *   - 8 propagator CTAs (32 digits each) consume AGG words in tile order, 128 tiles per
 *     round, and publish every tile's exclusive prefix (PREF) -- each word is read once
 *     and written once, whatever the backlog is;
 *   - a worker consumes PREF for tile i only after it has ranked and staged tile i+1, so
 *     the propagation latency is hidden behind useful work, and the keys of the next tile
 *     are loaded while the previous one is written out.
 * Tiles are handed out by an atomic ticket (in order, to running CTAs only), workers never
 * wait on each other, and a worker publishes AGG of its current tile before it waits for
 * PREF of its previous one: no circular wait.  AGG/PREF words are reset by their single
 * consumer, so the arrays are all-zero again when the kernel ends (no per-pass memset).
 *
 * Ranking / verification / repair are the same as in the classic kernel (see radix.cu).
 */
#ifndef CLO_RADIX_PP_CUH
#define CLO_RADIX_PP_CUH

#include "radix_prop.cuh"

template <typename ElemT, bool HAS_VAL, bool IDENTITY, typename LbT, int THREADS, int IPT, int RANK_MODE>
__global__ void __launch_bounds__(THREADS, (THREADS >= 512 ? 2 : 4))
clo_radix_onesweep_pp(const ElemT* __restrict__ in, ElemT* __restrict__ out,
		const u32* __restrict__ vin, u32* __restrict__ vout, size_t n, u32 num_tiles,
		LbT* __restrict__ agg, LbT* __restrict__ pref, u32* __restrict__ ticket,
		const u64* __restrict__ bins_base, u32 start_bit, u32 dmask, CloKeySpec ks,
		int* __restrict__ err_flag, int prof_on, int flags) {
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * IPT;
	constexpr bool VERIFY = (RANK_MODE == RANK_ATOMIC);
	constexpr bool USE_INFO = HAS_VAL || !IDENTITY;
	static_assert(THREADS >= RADIX, "one thread per digit is needed");
	static_assert(TILE <= 65536, "index-in-tile must fit 16 bits");

	extern __shared__ __align__(16) unsigned char smem_raw[];

	/* The propagators are latency critical and tiny; a worker sharing their SM would put its
	 * shared-memory traffic in front of every propagator load (measured: 2x slower chain).
	 * Each propagator publishes the id of its SM; a worker that finds itself on one of those
	 * SMs leaves (tiles are handed out by ticket, so nothing is lost). */
	u32* prop_sm = reinterpret_cast<u32*>(err_flag + 48);       /* [PP_NUM_PROP], zeroed per call */
	u32 my_sm;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(my_sm));
	if (blockIdx.x < PP_NUM_PROP) {
		if (threadIdx.x == 0) st_relaxed(prop_sm + blockIdx.x, my_sm + 1u);
		pp_propagate<LbT, THREADS>(agg, pref, num_tiles, err_flag, smem_raw, prof_on);
		return;
	}
	if (flags & 4) {
		__shared__ int s_leave;
		if (threadIdx.x == 0) {
			int leave = 0;
			for (int k = 0; k < PP_NUM_PROP; ++k) {
				u32 v = ld_relaxed(prop_sm + k);
				unsigned spins = 0;
				while (v == 0 && ++spins < (1u << 20)) v = ld_relaxed(prop_sm + k);
				if (v == my_sm + 1u) leave = 1;
			}
			s_leave = leave;
		}
		__syncthreads();
		if (s_leave) return;
	}
	/* optional phase profile (thread 0 of every worker): cycles per phase */
	u64* prof = reinterpret_cast<u64*>(err_flag + 16);
	long long t_prev = 0;
	auto mark = [&](int phase) {
		if (prof_on && threadIdx.x == 0) {
			const long long t = clock64();
			atomicAdd(prof + phase, (u64) (t - t_prev));
			t_prev = t;
		}
	};
	if (prof_on && threadIdx.x == 0) t_prev = clock64();

	/* ---- shared memory: warp histograms, two staged-tile buffers, per-buffer tables */
	constexpr size_t BUF_BYTES = (size_t) TILE * sizeof(ElemT) + (USE_INFO ? (size_t) TILE * 4 : 0) +
		(HAS_VAL ? (size_t) TILE * 4 : 0);
	u32* whist = reinterpret_cast<u32*>(smem_raw);                          /* [WARPS][RADIX] */
	u32* s_ds = whist + WARPS * RADIX;                                      /* [2][RADIX] digit starts */
	LbT* s_goff = reinterpret_cast<LbT*>(s_ds + 2 * RADIX);                 /* [2][RADIX], u64-sized slots */
	u32* s_misc = s_ds + 2 * RADIX + 4 * RADIX;                             /* [16] */
	unsigned char* s_buf = reinterpret_cast<unsigned char*>(s_misc + 16);   /* [2][BUF_BYTES] */

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const ElemT low_mask = (ElemT) ((((ElemT) dmask) << start_bit) | ((((ElemT) 1) << start_bit) - 1));
	const u32 wbase = (u32) warp * 32u * IPT + lane;
	u32* wh = whist + warp * RADIX;
	const LbT my_base = tid < RADIX ? (LbT) bins_base[tid] : (LbT) 0;   /* global start of my digit */

	ElemT key[IPT];
	u32 val[HAS_VAL ? IPT : 1];
	u32 pos2[(IPT + 1) / 2];

	auto tile_count_of = [&](u32 t) -> u32 {
		const size_t base = (size_t) t * TILE;
		return (base + TILE <= n) ? (u32) TILE : (u32) (n - base);
	};
	/* per-buffer offset table: slot b is a u64[RADIX]-sized region */
	auto goff_of = [&](int b) { return reinterpret_cast<LbT*>(reinterpret_cast<u64*>(s_goff) + (size_t) b * RADIX); };
	auto buf_keys = [&](int b) { return reinterpret_cast<ElemT*>(s_buf + (size_t) b * BUF_BYTES); };
	auto buf_info = [&](int b) { return reinterpret_cast<u32*>(s_buf + (size_t) b * BUF_BYTES + (size_t) TILE * sizeof(ElemT)); };
	auto buf_vals = [&](int b) {
		return reinterpret_cast<u32*>(s_buf + (size_t) b * BUF_BYTES + (size_t) TILE * sizeof(ElemT) + (USE_INFO ? (size_t) TILE * 4 : 0));
	};

	auto load_tile = [&](u32 t, bool streaming) {
		const size_t base = (size_t) t * TILE;
		const u32 cnt = tile_count_of(t);
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			if (cnt == (u32) TILE || local < cnt) {
				key[i] = streaming ? __ldcs(in + base + local) : in[base + local];
				if (HAS_VAL) val[i] = streaming ? __ldcs(vin + base + local) : vin[base + local];
			} else {
				key[i] = ElemT(0);
				if (HAS_VAL) val[i] = 0;
			}
		}
	};
	auto rank_ballot = [&](u32 cnt) {
#pragma unroll
		for (int i = 0; i < (IPT + 1) / 2; ++i) pos2[i] = 0;
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			const bool valid = local < cnt;
			const u32 d = radix_digit<ElemT, IDENTITY>(key[i], ks, start_bit, dmask);
			u32 peers = match_digit_ballot(d);
			peers &= __ballot_sync(0xffffffffu, valid);
			const u32 lt = peers & lanemask_lt();
			u32 old = 0;
			if (valid && lt == 0) {
				old = wh[d];
				wh[d] = old + __popc(peers);
			}
			__syncwarp();
			const int leader = __ffs(peers) - 1;
			old = __shfl_sync(0xffffffffu, old, leader & 31);
			pos2[i >> 1] |= (old + __popc(lt)) << (16 * (i & 1));
		}
	};
	auto rank_atomic = [&](auto full_tag, u32 cnt) {
		constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
		for (int i = 0; i < (IPT + 1) / 2; ++i) pos2[i] = 0;
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			if (FULL || local < cnt)
				pos2[i >> 1] |= atomicAdd(&wh[radix_digit<ElemT, IDENTITY>(key[i], ks, start_bit, dmask)], 1u) << (16 * (i & 1));
		}
	};
	auto digit_count = [&]() -> u32 {
		u32 count = 0;
		if (tid < RADIX) {
#pragma unroll
			for (int w = 0; w < WARPS; ++w) count += whist[w * RADIX + tid];
		}
		return count;
	};
	/* contains one barrier; leaves whist[w][d] = dstart[d] + count of d in lower warps */
	auto digit_starts = [&](u32 count, int b) {
		const u32 incl = warp_inclusive_scan<u32>(count, lane);
		if (tid < RADIX && lane == 31) s_misc[1 + warp] = incl;
		__syncthreads();
		if (tid < RADIX) {
			u32 off = 0;
#pragma unroll
			for (int w = 0; w < RADIX / 32; ++w) if (w < warp) off += s_misc[1 + w];
			const u32 ds = off + incl - count;
			s_ds[b * RADIX + tid] = ds;
			u32 c[WARPS];
#pragma unroll
			for (int w = 0; w < WARPS; ++w) c[w] = whist[w * RADIX + tid];
			u32 run = ds;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) { whist[w * RADIX + tid] = run; run += c[w]; }
		}
	};
	auto stage = [&](auto full_tag, u32 cnt, int b) {
		constexpr bool FULL = decltype(full_tag)::value;
		ElemT* skeys = buf_keys(b);
		u32* sinfo = buf_info(b);
		u32* svals = buf_vals(b);
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			if (FULL || local < cnt) {
				const u32 d = radix_digit<ElemT, IDENTITY>(key[i], ks, start_bit, dmask);
				const u32 p = wh[d] + ((pos2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
				skeys[p] = key[i];
				if (USE_INFO) sinfo[p] = (d << 16) | local;
				if (HAS_VAL) svals[p] = val[i];
			}
		}
	};
	auto write_out = [&](auto full_tag, u32 cnt, int b, bool verify) -> bool {
		constexpr bool FULL = decltype(full_tag)::value;
		const ElemT* skeys = buf_keys(b);
		const u32* sinfo = buf_info(b);
		const u32* svals = buf_vals(b);
		const LbT* goff = goff_of(b);
		bool bad = false;
#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 j = (u32) tid + i * THREADS;
			if (FULL || j < cnt) {
				const ElemT k = skeys[j];
				u32 d;
				if (USE_INFO) {
					const u32 info = sinfo[j];
					if (verify && j > 0 && info <= sinfo[j - 1]) bad = true;
					d = info >> 16;
				} else {
					if (verify && j > 0 && (k & low_mask) < (skeys[j - 1] & low_mask)) bad = true;
					d = radix_digit<ElemT, IDENTITY>(k, ks, start_bit, dmask);
				}
				const LbT o = goff[d] + (LbT) j;
				out[o] = k;
				if (HAS_VAL) vout[o] = svals[j];
			}
		}
		return bad;
	};
	/* front half of a tile: rank, publish the digit counts, stage into buffer b */
	auto front = [&](u32 t, int b) {
		const u32 cnt = tile_count_of(t);
		const bool full = cnt == (u32) TILE;
		for (int i = tid; i < WARPS * RADIX; i += THREADS) whist[i] = 0;
		__syncthreads();
		mark(0);                                 /* zero + (first use of key[]: wait for the loads) */
		if (RANK_MODE == RANK_ATOMIC) { if (full) rank_atomic(std::true_type{}, cnt); else rank_atomic(std::false_type{}, cnt); }
		else rank_ballot(cnt);
		__syncthreads();
		mark(1);                                 /* rank */
		const u32 count = digit_count();
		if (tid < RADIX)
			st_relaxed(agg + (size_t) t * RADIX + tid, (LbT) (PPWord<LbT>::VALID | (LbT) count));
		digit_starts(count, b);
		__syncthreads();
		mark(2);                                 /* digit phase */
		if (full) stage(std::true_type{}, cnt, b); else stage(std::false_type{}, cnt, b);
		mark(3);                                 /* stage */
	};
	/* back half, part 1: wait for the tile's exclusive prefix, build the offset table */
	auto back_wait = [&](u32 t, int b, LbT w) {
		LbT* goff = goff_of(b);
		if (tid < RADIX) {
			LbT* p = pref + (size_t) t * RADIX + tid;
			unsigned spins = 0;
			while (!(w & PPWord<LbT>::VALID)) {
				if (++spins > SPIN_LIMIT) { atomicExch(err_flag, 1); break; }
				w = ld_relaxed(p);
			}
			st_relaxed(p, (LbT) 0);                                  /* consumed: reset */
			goff[tid] = my_base + (w & PPWord<LbT>::VAL) - (LbT) s_ds[b * RADIX + tid];
			if (prof_on && tid == 0) atomicAdd(prof + 8, (u64) spins);   /* PREF polls that found nothing */
		}
		__syncthreads();      /* goff ready; staging of this buffer finished long ago */
		mark(5);                                 /* prefix wait */
	};
	/* back half, part 2: write the staged tile out (verify; repair if needed) */
	auto back_write = [&](u32 t, int b) -> bool {
		const u32 cnt = tile_count_of(t);
		const bool full = cnt == (u32) TILE;
		const bool bad = full ? write_out(std::true_type{}, cnt, b, VERIFY) : write_out(std::false_type{}, cnt, b, VERIFY);
		mark(6);                                 /* write-out */
		if (VERIFY) {
			if (__syncthreads_or(bad ? 1 : 0)) {
				/* atomic ranks were not in lane order somewhere in this tile: redo it with the
				 * ballot ranks (counts and offsets are unchanged).  Clobbers key[]/whist: the
				 * caller re-establishes them. */
				if (tid == 0) atomicAdd(err_flag + 1, 1);
				for (int i = tid; i < WARPS * RADIX; i += THREADS) whist[i] = 0;
				__syncthreads();
				load_tile(t, false);
				rank_ballot(cnt);
				__syncthreads();
				const u32 count2 = digit_count();
				digit_starts(count2, b);
				__syncthreads();
				stage(std::false_type{}, cnt, b);
				__syncthreads();
				write_out(std::false_type{}, cnt, b, false);
				__syncthreads();
				return true;
			}
		}
		return false;
	};
	auto take_ticket = [&]() -> u32 {
		/* two barriers: every thread has read s_misc[0] before it can be overwritten again */
		if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
		__syncthreads();
		const u32 t = s_misc[0];
		__syncthreads();
		return t;
	};

	/* ---- pipeline: front(cur) ; load(next) ; back(prev).  Everything with a long latency is
	 *      requested one step early: the ticket of the tile after next, the keys of the next
	 *      tile, the prefix word of the previous tile. */
	/* ---- pipeline per iteration:
	 *        front(cur)          rank, publish AGG[cur], stage
	 *        back_wait(prev)     PREF[prev] (requested before front) -> offsets
	 *        ticket + load(next) drawn as late as possible: the propagation is in tile order, so
	 *                            a ticket that is held long before its AGG is published delays
	 *                            every later tile; the loads still overlap the write-out
	 *        back_write(prev)    coalesced write-out of the previous tile */
	const bool pf_pref = (flags & 2) == 0;       /* request the prefix word before front() (default on) */
	u32 cur = take_ticket();
	if (cur >= num_tiles) return;
	load_tile(cur, true);
	u32 prev = 0xffffffffu;
	int b = 0;
	for (;;) {
		LbT wp = 0;
		if (pf_pref && prev != 0xffffffffu && tid < RADIX) wp = ld_relaxed(pref + (size_t) prev * RADIX + tid);
		front(cur, b);
		if (prev != 0xffffffffu) back_wait(prev, b ^ 1, wp);
		const u32 nxt = take_ticket();
		const bool more = nxt < num_tiles;
		if (more) load_tile(nxt, true);        /* in flight while the previous tile is written out */
		mark(4);
		if (prev != 0xffffffffu) {
			if (back_write(prev, b ^ 1) && more) load_tile(nxt, true);   /* repair clobbered key[] */
		}
		prev = cur;
		b ^= 1;
		if (!more) break;
		cur = nxt;
	}
	back_wait(prev, b ^ 1, (LbT) 0);
	back_write(prev, b ^ 1);
}

template <typename ElemT, bool HAS_VAL, bool IDENTITY, int THREADS, int IPT, typename LbT>
constexpr size_t onesweep_pp_smem() {
	constexpr bool USE_INFO = HAS_VAL || !IDENTITY;
	constexpr size_t worker = (size_t) (THREADS / 32) * RADIX * 4 + 2 * RADIX * 4 + 4 * RADIX * 4 + 16 * 4 +
		2 * ((size_t) THREADS * IPT * sizeof(ElemT) + (USE_INFO ? (size_t) THREADS * IPT * 4 : 0) +
			(HAS_VAL ? (size_t) THREADS * IPT * 4 : 0));
	constexpr size_t prop = ((size_t) PP_WINDOW * 32 + 8 * 32) * sizeof(LbT) + 16;
	return worker > prop ? worker : prop;
}

#endif
