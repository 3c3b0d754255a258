/*
 * radix_prop.cuh -- what every persistent onesweep kernel shares: the radix constants, the
 * ballot digit match and the PREFIX-PROPAGATOR CTAs with their AGG/PREF word protocol.
 * Included (inside an anonymous namespace) by radix.cu and radix_v6.cu.
 */
#ifndef CLO_RADIX_PROP_CUH
#define CLO_RADIX_PROP_CUH

const int RADIX_BITS = 8;
const int RADIX = 1 << RADIX_BITS;
const unsigned SPIN_LIMIT = 1u << 24;

/* lanes of the warp whose digit equals mine: one ballot per digit bit, 4 SASS
 * instructions per bit (test, vote, conditional invert, and) */
__device__ __forceinline__ u32 match_digit_ballot(u32 d) {
	u32 peers = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < RADIX_BITS; ++b) {
		asm("{\n\t"
			".reg .pred p;\n\t"
			".reg .b32 m, t;\n\t"
			"and.b32 t, %1, %2;\n\t"
			"setp.ne.u32 p, t, 0;\n\t"
			"vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
			"@!p not.b32 m, m;\n\t"
			"and.b32 %0, %0, m;\n\t"
			"}" : "+r"(peers) : "r"(d), "r"(1u << b));
	}
	return peers;
}

const int PP_NUM_PROP = 8;    /* propagator CTAs: RADIX / 32 digits each */
const int PP_WINDOW = 128;    /* tiles a propagator looks at per iteration */

template <typename LbT> struct PPWord;
template <> struct PPWord<u32> { static constexpr u32 VALID = 1u << 31, VAL = (1u << 31) - 1; };
template <> struct PPWord<u64> { static constexpr u64 VALID = 1ull << 63, VAL = (1ull << 63) - 1; };

/* In-order streaming propagation for 32 digits (one per lane).  Invariant at the top of an
 * iteration: AGG of tiles < t0 is consumed (and reset), PREF of tiles <= t0 is published.
 * An iteration reads the window AGG[t0 .. t0+R), takes its longest valid prefix L, consumes
 * those L tiles and publishes PREF[t0+1 .. t0+L].  PREF[t] therefore depends on AGG[< t]
 * only -- never on a later tile -- which is what makes the worker pipeline deadlock free. */
template <typename LbT, int THREADS>
__device__ __forceinline__ void pp_propagate(LbT* __restrict__ agg, LbT* __restrict__ pref, u32 num_tiles,
		int* __restrict__ err_flag, unsigned char* smem_raw, int prof_on) {
	constexpr int WARPS = THREADS / 32;
	constexpr int PP_U = PP_WINDOW / WARPS;        /* window entries per thread */
	constexpr int R = PP_WINDOW;                   /* window: tiles per iteration */
	constexpr int SEG = 4;                         /* warps doing the serial part */
	static_assert(R % SEG == 0, "window must split into segments");
	LbT* s_val = reinterpret_cast<LbT*>(smem_raw);         /* [R][32] */
	LbT* s_seg = s_val + R * 32;                            /* [SEG][32] */
	int* s_first = reinterpret_cast<int*>(s_seg + SEG * 32);   /* [2] first unpublished index (double buffered) */
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const u32 d = blockIdx.x * 32u + lane;
	LbT running = 0;
	unsigned idle = 0;
	u64* prof = reinterpret_cast<u64*>(err_flag + 16);
	const long long p_t0 = prof_on ? clock64() : 0;
	unsigned long long rounds = 0;
	if (threadIdx.x < 2) s_first[threadIdx.x] = R;
	if (warp == 0) st_relaxed(pref + d, (LbT) PPWord<LbT>::VALID);      /* PREF[0] = 0 */
	__syncthreads();
	u32 t0 = 0;
	int par = 0;
	/* window entry j = u * WARPS + warp is tile t0 + j */
	LbT w[PP_U];
	auto load_window = [&](u32 base) {
#pragma unroll
		for (int u = 0; u < PP_U; ++u) {
			const u32 t = base + u * WARPS + warp;
			w[u] = (t < num_tiles) ? ld_relaxed(agg + (size_t) t * RADIX + d) : (LbT) 0;
		}
	};
	load_window(0);
	while (t0 < num_tiles) {
		int first_bad = R;
#pragma unroll
		for (int u = PP_U - 1; u >= 0; --u) {
			const int j = u * WARPS + warp;
			if (!(w[u] & PPWord<LbT>::VALID)) first_bad = j;
			s_val[j * 32 + lane] = w[u] & PPWord<LbT>::VAL;
		}
		/* a tile counts only when all 32 digit words of it are there */
		first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, 16));
		first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, 8));
		first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, 4));
		first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, 2));
		first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, 1));
		if (lane == 0 && first_bad < R) atomicMin(&s_first[par], first_bad);
		__syncthreads();
		const int L = s_first[par];                /* tiles t0 .. t0+L-1 are complete */
		if (threadIdx.x == 0) s_first[par ^ 1] = R;   /* reset the other slot for the next iteration */
		/* The next window starts at t0 + L.  Its loads are issued NOW, before this
		 * iteration's scan and stores, so that one L2 round trip overlaps the other work (and
		 * is not ordered behind the strong stores below).  s_val already holds this window. */
		load_window(t0 + (u32) L);
		if (L == 0) {
			if (++idle > (SPIN_LIMIT >> 3)) { atomicExch(err_flag, 1); break; }
			par ^= 1;
			__syncthreads();
			continue;
		}
		/* inclusive prefix along j for every digit: SEG warps scan R/SEG entries each
		 * (all loads first, so the shared-memory latency is paid once, not per step) */
		if (warp < SEG) {
			constexpr int K = R / SEG;
			LbT v[K];
#pragma unroll
			for (int k = 0; k < K; ++k) v[k] = s_val[(warp * K + k) * 32 + lane];
#pragma unroll
			for (int k = 1; k < K; ++k) v[k] += v[k - 1];
#pragma unroll
			for (int k = 0; k < K; ++k) s_val[(warp * K + k) * 32 + lane] = v[k];
			s_seg[warp * 32 + lane] = v[K - 1];
		}
		__syncthreads();
		LbT seg_off[SEG];
		{
			LbT run = 0;
#pragma unroll
			for (int q = 0; q < SEG; ++q) { seg_off[q] = run; run += s_seg[q * 32 + lane]; }
		}
		LbT consumed_total = 0;
		{
			const int jl = L - 1;                  /* inclusive prefix through the last consumed tile */
			LbT off = 0;
#pragma unroll
			for (int q = 0; q < SEG; ++q) if (jl / (R / SEG) == q) off = seg_off[q];
			consumed_total = off + s_val[jl * 32 + lane];
		}
#pragma unroll
		for (int u = 0; u < PP_U; ++u) {
			const int j = u * WARPS + warp;
			if (j < L) {
				const u32 t = t0 + j;
				LbT off = 0;
#pragma unroll
				for (int q = 0; q < SEG; ++q) if (j / (R / SEG) == q) off = seg_off[q];
				st_relaxed(agg + (size_t) t * RADIX + d, (LbT) 0);                     /* consumed: reset */
				if (t + 1 < num_tiles)
					st_relaxed(pref + (size_t) (t + 1) * RADIX + d,
						(LbT) (PPWord<LbT>::VALID | ((running + off + s_val[j * 32 + lane]) & PPWord<LbT>::VAL)));
			}
		}
		running += consumed_total;
		t0 += (u32) L;
		par ^= 1;
		idle = 0;
		++rounds;
		__syncthreads();
	}
	if (prof_on && threadIdx.x == 0 && blockIdx.x == 0) {
		atomicAdd(prof + 12, (u64) (clock64() - p_t0));     /* propagator 0: total cycles */
		atomicAdd(prof + 13, (u64) rounds);                 /* ... and productive iterations */
	}
}


/* Second-generation propagator: G independent groups per CTA (one named barrier each), each
 * group owns 32 digits (lane = digit) and looks at a window of PP2_WINDOW tiles per round;
 * warp wg of the group owns a CONTIGUOUS chunk of C = window / warps-per-group tiles, scans it
 * in registers and publishes only its chunk total -- no window-wide shared-memory scan, one
 * barrier per round.  Same protocol and invariant as pp_propagate:
 * at the top of a round AGG of tiles < t0 is consumed (reset), PREF of tiles <= t0 published. */
const int PP2_CHUNK = 16;                   /* tiles per warp and round; window = warps per group x 16 */

template <typename LbT, int THREADS, int G>
__device__ __forceinline__ void pp_propagate2(LbT* __restrict__ agg, LbT* __restrict__ pref, u32 num_tiles,
		int* __restrict__ err_flag, unsigned char* smem_raw, int prof_on, int cta_index) {
	constexpr int WARPS = THREADS / 32;
	constexpr int WPG = WARPS / G;                 /* warps per group */
	constexpr int C = sizeof(LbT) == 8 ? PP2_CHUNK / 2 : PP2_CHUNK;   /* tiles per warp and round (64-bit words: half, to stay in registers) */
	static_assert(WARPS % G == 0, "bad propagator shape");
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int grp = warp / WPG, wg = warp % WPG;
	/* shared: per group, double buffered: fb[WPG] ints + tot[WPG][32] LbT */
	LbT* s_tot = reinterpret_cast<LbT*>(smem_raw) + (size_t) grp * 2 * WPG * 32;          /* [2][WPG][32] */
	int* s_fb = reinterpret_cast<int*>(reinterpret_cast<LbT*>(smem_raw) + (size_t) G * 2 * WPG * 32) + grp * 2 * WPG;   /* [2][WPG] */
	const u32 d = (u32) (cta_index * G + grp) * 32u + lane;
	const int bar_id = 1 + grp;
	LbT running = 0;
	unsigned idle = 0;
	u64* prof = reinterpret_cast<u64*>(err_flag + 16);
	const long long p_t0 = prof_on ? clock64() : 0;
	unsigned long long rounds = 0;
	if (wg == 0) st_relaxed(pref + d, (LbT) PPWord<LbT>::VALID);      /* PREF[0] = 0 */
	u32 t0 = 0;
	int par = 0;
	LbT w[C];
	auto load_window = [&](u32 base) {
#pragma unroll
		for (int u = 0; u < C; ++u) {
			const u32 t = base + (u32) (wg * C + u);
			w[u] = (t < num_tiles) ? ld_relaxed(agg + (size_t) t * RADIX + d) : (LbT) 0;
		}
	};
	load_window(0);
	while (t0 < num_tiles) {
		/* my chunk: first entry that is not complete (all 32 digit words there), inclusive scan */
		int fb = C;
		LbT v[C];
#pragma unroll
		for (int u = C - 1; u >= 0; --u)
			if (!__all_sync(0xffffffffu, (w[u] & PPWord<LbT>::VALID) != 0)) fb = u;
		LbT acc = 0, tot = 0;
#pragma unroll
		for (int u = 0; u < C; ++u) {
			acc += w[u] & PPWord<LbT>::VAL;
			v[u] = acc;
			if (u < fb) tot = acc;
		}
		s_tot[(par * WPG + wg) * 32 + lane] = tot;
		if (lane == 0) s_fb[par * WPG + wg] = fb;
		asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(WPG * 32) : "memory");
		/* consumed tiles L = full chunks + the valid part of the first incomplete one */
		int L = 0;
		LbT off = 0, consumed = 0;
		bool open = true;
#pragma unroll
		for (int q = 0; q < WPG; ++q) {
			const int fq = s_fb[par * WPG + q];
			const LbT tq = s_tot[(par * WPG + q) * 32 + lane];
			if (open) {
				L += fq;
				consumed += tq;
				if (q < wg) off += tq;
				if (fq < C) open = false;
			}
		}
		const int my_n = min(max(L - wg * C, 0), C);
		const u32 my_t = t0 + (u32) (wg * C);
		/* next window: issued before the stores so that one L2 round trip overlaps them */
		load_window(t0 + (u32) L);
		if (L == 0) {
			if (++idle > (SPIN_LIMIT >> 3)) { atomicExch(err_flag, 1); break; }
			par ^= 1;
			continue;
		}
#pragma unroll
		for (int u = 0; u < C; ++u) {
			if (u < my_n) {
				const u32 t = my_t + u;
				st_relaxed(agg + (size_t) t * RADIX + d, (LbT) 0);                         /* consumed: reset */
				if (t + 1 < num_tiles)
					st_relaxed(pref + (size_t) (t + 1) * RADIX + d,
						(LbT) (PPWord<LbT>::VALID | ((running + off + v[u]) & PPWord<LbT>::VAL)));
			}
		}
		running += consumed;
		t0 += (u32) L;
		par ^= 1;
		idle = 0;
		++rounds;
	}
	if (prof_on && threadIdx.x == 0 && cta_index == 0) {
		atomicAdd(prof + 12, (u64) (clock64() - p_t0));
		atomicAdd(prof + 13, (u64) rounds);
	}
}

template <typename LbT, int THREADS, int G>
constexpr size_t pp_propagate2_smem() {
	return (size_t) G * 2 * (THREADS / 32 / G) * 32 * sizeof(LbT) + (size_t) G * 2 * (THREADS / 32 / G) * 4 + 16;
}

#endif
