/*
 * radix.cu -- LSD radix sort as a ONESWEEP: one histogram pass over the keys,
 * then one pass per 8-bit digit that ranks, looks back and scatters in a single
 * kernel.  Replaces the reference's per-digit {localsort, histogram, scan,
 * scatter} sequence (/root/reference/src/cl_ops/sort/clo_sort_satradix.cl:34-258,
 * host loop clo_sort_satradix.c:264-313) while keeping its result: a stable
 * ascending sort on the raw key bits.
 *
 * Per pass a CTA
 *   1. takes a tile ticket (tiles are handed out in launch order, which keeps the
 *      look-back deadlock free and the sort stable),
 *   2. loads TILE keys warp-striped (coalesced 128 B per warp instruction),
 *   3. ranks every key inside its warp with ballot-built digit-match masks and a
 *      warp-private shared-memory histogram (no atomics),
 *   4. turns the warp histograms into tile-local digit offsets and publishes the
 *      tile's 256 digit counts; decoupled look-back over the predecessor tiles
 *      gives each digit's global offset (no global scan kernel),
 *   5. stages the tile in shared memory in digit order and writes it out so that
 *      consecutive threads write consecutive addresses within each digit run.
 *
 * HBM traffic (u32 keys): 4 B/key for the histogram + 4 passes x 8 B/key = 36 B/key.
 */
#include "clo_internal.h"
#include "device_utils.cuh"
#include "sort_common.h"

#include <cstring>
#include <type_traits>

using namespace clo;

namespace {

int g_radix_profile = 0;
int g_pp_flags = 0;

#include "radix_prop.cuh"
const int MAX_PASSES = 8;
#ifndef CLO_IPT_U32
#define CLO_IPT_U32 16
#endif
const int LB_FIRST = 4;     /* look-back window: first load batch */
const int LB_NEXT = 4;      /* ... and the following ones */

struct PassCfg {
	int passes;
	u32 start_bit[MAX_PASSES];
	u32 dmask[MAX_PASSES];
};

/* look-back word: 2 flag bits on top of the value */
template <typename LbT> struct Lb;
template <> struct Lb<u32> {
	static constexpr u32 AGG = 1u << 30, PREFIX = 2u << 30, VAL = (1u << 30) - 1, FLAGS = 3u << 30;
};
template <> struct Lb<u64> {
	static constexpr u64 AGG = 1ull << 62, PREFIX = 2ull << 62, VAL = (1ull << 62) - 1, FLAGS = 3ull << 62;
};

/* raw element -> (promoted) key bits the digits are cut from */
template <typename ElemT, bool IDENTITY>
__device__ __forceinline__ u64 radix_key(ElemT e, const CloKeySpec& ks) {
	if (IDENTITY) return (u64) e;
	u64 k = clo_extract_key((u64) e, ks);
	/* OpenCL C promotes a sub-int signed key to int before `>>`
	 * (clo_sort_satradix.cl:61): sign-extend to 32 bits */
	if (ks.key_kind == CLO_KIND_SIGNED && ks.key_bits < 32) {
		if (k & (1ull << (ks.key_bits - 1))) k |= (0xffffffffull & ~((1ull << ks.key_bits) - 1));
	}
	return k;
}

template <typename ElemT, bool IDENTITY>
__device__ __forceinline__ u32 radix_digit(ElemT e, const CloKeySpec& ks, u32 start_bit, u32 dmask) {
	if (IDENTITY) {
		if (sizeof(ElemT) <= 4) return (((u32) e) >> start_bit) & dmask;
		return (u32) (((u64) e) >> start_bit) & dmask;
	}
	return (u32) (radix_key<ElemT, false>(e, ks) >> start_bit) & dmask;
}

/* ------------------------------------------------------------- histogram */

/* One read of the keys -> the 256-bin histogram of every digit.
 * Shared-memory layout [pass][bin][COLS]: lane l only ever touches column l % COLS, so
 * with COLS == 32 the 32 atomics of a warp instruction hit 32 different banks no matter
 * what the digits are (uniform or constant keys cost the same: one wavefront).  The
 * atomics are still needed because the warps of the CTA share the table. */
template <typename ElemT, bool IDENTITY, int PASSES, int COLS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
clo_radix_histogram(const ElemT* __restrict__ in, size_t n, u64* __restrict__ ghist,
		PassCfg cfg, CloKeySpec ks, int vec_ok) {
	extern __shared__ __align__(16) unsigned char smem_raw[];
	u32* sh = reinterpret_cast<u32*>(smem_raw);          /* [PASSES][RADIX][COLS] */
	for (int i = threadIdx.x; i < PASSES * RADIX * COLS; i += THREADS) sh[i] = 0;
	__syncthreads();
	const u32 col = threadIdx.x & (COLS - 1);
	u32 sb[PASSES], dm[PASSES];
#pragma unroll
	for (int p = 0; p < PASSES; ++p) { sb[p] = cfg.start_bit[p]; dm[p] = cfg.dmask[p]; }
	auto count = [&](ElemT e) {
		const u64 k = radix_key<ElemT, IDENTITY>(e, ks);
#pragma unroll
		for (int p = 0; p < PASSES; ++p) {
			const u32 d = (sizeof(ElemT) <= 4 && IDENTITY) ? (((u32) k >> sb[p]) & dm[p]) : ((u32) (k >> sb[p]) & dm[p]);
			atomicAdd(&sh[(p * RADIX + d) * COLS + col], 1u);
		}
	};
	constexpr int EPV = 16 / sizeof(ElemT);
	const size_t stride = (size_t) gridDim.x * THREADS;
	const size_t tid = (size_t) blockIdx.x * THREADS + threadIdx.x;
	if (vec_ok) {
		const size_t nvec = n / EPV;
		size_t i = tid;
		/* two vectors in flight per thread */
		for (; i + stride < nvec; i += 2 * stride) {
			ElemT e0[EPV], e1[EPV];
			load_vec_cs<ElemT, EPV>(in + i * EPV, e0);
			load_vec_cs<ElemT, EPV>(in + (i + stride) * EPV, e1);
#pragma unroll
			for (int c = 0; c < EPV; ++c) count(e0[c]);
#pragma unroll
			for (int c = 0; c < EPV; ++c) count(e1[c]);
		}
		for (; i < nvec; i += stride) {
			ElemT e0[EPV];
			load_vec_cs<ElemT, EPV>(in + i * EPV, e0);
#pragma unroll
			for (int c = 0; c < EPV; ++c) count(e0[c]);
		}
		for (size_t r = nvec * EPV + tid; r < n; r += stride) count(in[r]);
	} else {
		for (size_t r = tid; r < n; r += stride) count(in[r]);
	}
	__syncthreads();
	/* fold the columns (rotated start: conflict free) and add to the global histogram */
	for (int b = threadIdx.x; b < PASSES * RADIX; b += THREADS) {
		u32 t = 0;
#pragma unroll 8
		for (int c = 0; c < COLS; ++c) t += sh[b * COLS + ((c + threadIdx.x) & (COLS - 1))];
		if (t) atomicAdd(&ghist[b], (u64) t);
	}
}

/* exclusive scan of each pass's 256 bins: block p handles pass p */
__global__ void __launch_bounds__(RADIX)
clo_radix_scan_bins(const u64* __restrict__ ghist, u64* __restrict__ bins_base) {
	__shared__ u64 s_w[RADIX / 32];
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	const u64 c = ghist[blockIdx.x * RADIX + t];
	const u64 incl = warp_inclusive_scan<u64>(c, lane);
	if (lane == 31) s_w[warp] = incl;
	__syncthreads();
	u64 off = 0;
	for (int w = 0; w < warp; ++w) off += s_w[w];
	bins_base[blockIdx.x * RADIX + t] = off + incl - c;
}

/* -------------------------------------------------------------- onesweep */

/* A digit functor for the sample-sort partition: bucket = number of splitters
 * (key, global index) that are <= (my key, my global index). */
struct SplitterArgs {
	const void* keys;       /* nparts-1 splitter keys, element type */
	const u64* idx;         /* nparts-1 splitter global indices */
	u32 count;              /* nparts-1 */
	u64 gidx0;              /* global index of element 0 */
};

/* How a key gets its rank among the equal-digit keys of its warp:
 *  RANK_BALLOT  digit-match masks from 8 ballots + a leader's read-modify-write of
 *               the warp's shared histogram.  ~45 ALU-pipe instructions per 32 keys:
 *               on B200 this alone costs 0.4 ms per 2^28-key pass (tools/ubench_match.cu).
 *  RANK_ATOMIC  one shared-memory atomicAdd per key on the warp-private histogram.
 *               The returned value is the stable rank iff the lanes of one warp
 *               instruction that hit the same address are served in lane order.  B200
 *               does that (0 violations in 3e8 samples, tools/ubench_match.cu) but it is
 *               not an architectural guarantee, so the kernel does not trust it: each
 *               staged tile is VERIFIED -- (digit << 16 | index-in-tile) must be strictly
 *               increasing along the staged order, which holds iff the tile-local
 *               partition is stable -- and a tile that fails is re-ranked with
 *               RANK_BALLOT and rewritten before the kernel moves on.  Correctness never
 *               depends on the atomic order; only speed does. */
enum { RANK_BALLOT = 0, RANK_ATOMIC = 1 };

template <typename ElemT, bool HAS_VAL, bool IDENTITY, bool PARTITION, typename LbT,
	int THREADS, int IPT, int RANK_MODE, int LBF = LB_FIRST, int LBN = LB_NEXT>
__global__ void __launch_bounds__(THREADS, (THREADS >= 384 ? 2 : 4))
clo_radix_onesweep(const ElemT* __restrict__ in, ElemT* __restrict__ out,
		const u32* __restrict__ vin, u32* __restrict__ vout, size_t n,
		LbT* __restrict__ lookback, u32* __restrict__ ticket, const u64* __restrict__ bins_base,
		u32 start_bit, u32 dmask, CloKeySpec ks, SplitterArgs sp, int* __restrict__ err_flag, int prof_on) {
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * IPT;
	constexpr bool VERIFY = (RANK_MODE == RANK_ATOMIC);
	/* the (digit, index) word is needed whenever the key alone cannot prove stability or
	 * give the digit back: payloads, extracted keys, the splitter partition */
	constexpr bool USE_INFO = HAS_VAL || !IDENTITY || PARTITION;
	static_assert(THREADS >= RADIX, "one thread per digit is needed");
	static_assert(TILE <= 65536, "index-in-tile must fit 16 bits");

	extern __shared__ __align__(16) unsigned char smem_raw[];
	u32* whist = reinterpret_cast<u32*>(smem_raw);                         /* [WARPS][RADIX] */
	u32* s_dstart = whist + WARPS * RADIX;                                 /* [RADIX] */
	LbT* s_goff = reinterpret_cast<LbT*>(s_dstart + RADIX);                /* [RADIX] (u64-sized slot) */
	u32* s_misc = s_dstart + RADIX + 2 * RADIX;                            /* [16]: tile, warp sums, flags */
	ElemT* skeys = reinterpret_cast<ElemT*>(s_misc + 16);                  /* [TILE] */
	u32* sinfo = reinterpret_cast<u32*>(skeys + TILE);                     /* [TILE] digit<<16 | index, if USE_INFO */
	u32* svals = sinfo + TILE;                                             /* [TILE] if HAS_VAL (then USE_INFO) */

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	/* bits of this digit and below (keys-only verification) */
	const ElemT low_mask = (ElemT) ((((ElemT) dmask) << start_bit) | ((((ElemT) 1) << start_bit) - 1));

	/* optional phase profile (CLO_RADIX_PROFILE=1): thread 0 accumulates cycles per phase
	 * into the u64 counters behind the status words */
	long long t_prev = 0;
	u64* prof = reinterpret_cast<u64*>(err_flag + 16);
	auto mark = [&](int phase) {
		if (prof_on && tid == 0) {
			const long long t = clock64();
			atomicAdd(prof + phase, (u64) (t - t_prev));
			t_prev = t;
		}
	};
	if (prof_on && tid == 0) t_prev = clock64();

	if (tid == 0) s_misc[0] = atomicAdd(ticket, 1u);
	for (int i = tid; i < WARPS * RADIX; i += THREADS) whist[i] = 0;
	__syncthreads();
	mark(0);
	const u32 tile = s_misc[0];
	const size_t tile_base = (size_t) tile * TILE;
	const bool full = tile_base + TILE <= n;
	const u32 tile_count = full ? (u32) TILE : (u32) (n - tile_base);

	/* The whole tile body is instantiated twice, for full tiles (no per-item bounds
	 * checks at all) and for the single partial tile at the end. */
	auto body = [&](auto full_tag) {
		constexpr bool FULL = decltype(full_tag)::value;
		/* ---- load, warp-striped: item i of lane l is element warp*32*IPT + i*32 + l */
		ElemT key[IPT];
		u32 val[HAS_VAL ? IPT : 1];
		u32 pos2[(IPT + 1) / 2];   /* ranks, two 16-bit halves per register */
#pragma unroll
		for (int i = 0; i < (IPT + 1) / 2; ++i) pos2[i] = 0;
		const u32 wbase = (u32) warp * 32u * IPT + lane;
	#pragma unroll
		for (int i = 0; i < IPT; ++i) {
			const u32 local = wbase + i * 32u;
			if (FULL || local < tile_count) {
				key[i] = __ldcs(in + tile_base + local);
				if (HAS_VAL) val[i] = __ldcs(vin + tile_base + local);
			} else {
				key[i] = ElemT(0);
				if (HAS_VAL) val[i] = 0;
			}
		}

		/* splitter table for the partition variant (tiny: <= 15 entries) */
		ElemT sp_key[PARTITION ? 15 : 1];
		u64 sp_idx[PARTITION ? 15 : 1];
		if (PARTITION) {
	#pragma unroll
			for (int s = 0; s < 15; ++s) {
				if (s < (int) sp.count) {
					sp_key[s] = reinterpret_cast<const ElemT*>(sp.keys)[s];
					sp_idx[s] = sp.idx[s];
				} else { sp_key[s] = ElemT(0); sp_idx[s] = 0; }
			}
		}

		auto digit_of = [&](ElemT k, u32 local) -> u32 {
			if (PARTITION) {
				const u64 g = sp.gidx0 + tile_base + local;
				u32 b = 0;
	#pragma unroll
				for (int s = 0; s < 15; ++s)
					if (s < (int) sp.count && (sp_key[s] < k || (sp_key[s] == k && sp_idx[s] <= g))) ++b;
				return b;
			}
			return radix_digit<ElemT, IDENTITY>(k, ks, start_bit, dmask);
		};

		u32* wh = whist + warp * RADIX;

		/* ---- rank inside the warp: pos[i] = number of earlier keys of this warp with my digit;
		 *      afterwards wh[d] = number of keys of this warp with digit d */
		auto rank_ballot = [&]() {
	#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 local = wbase + i * 32u;
				const bool valid = FULL || local < tile_count;
				const u32 d = digit_of(key[i], local);
				u32 peers = match_digit_ballot(d);
				if (!FULL) peers &= __ballot_sync(0xffffffffu, valid);
				const u32 lt = peers & lanemask_lt();
				u32 old = 0;
				if (valid && lt == 0) {            /* first lane of its digit group */
					old = wh[d];
					wh[d] = old + __popc(peers);
				}
				__syncwarp();
				const int leader = __ffs(peers) - 1;
				old = __shfl_sync(0xffffffffu, old, leader & 31);
				pos2[i >> 1] |= (old + __popc(lt)) << (16 * (i & 1));
			}
		};
		auto rank_atomic = [&]() {
	#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 local = wbase + i * 32u;
				if (FULL || local < tile_count)
					pos2[i >> 1] |= atomicAdd(&wh[digit_of(key[i], local)], 1u) << (16 * (i & 1));
			}
		};

		/* ---- per digit (thread d): tile count = sum of the warps' counts (independent loads) */
		auto digit_count = [&]() -> u32 {
			u32 count = 0;
			if (tid < RADIX) {
#pragma unroll
				for (int w = 0; w < WARPS; ++w) count += whist[w * RADIX + tid];
			}
			return count;
		};
		/* exclusive scan of the 256 counts -> start of each digit inside the tile; the per-warp
		 * counts become dstart[d] + (count of d in lower warps), so staging needs one lookup.
		 * Contains a barrier. */
		auto digit_starts = [&](u32 count) {
			const u32 incl = warp_inclusive_scan<u32>(count, lane);
			if (tid < RADIX && lane == 31) s_misc[1 + warp] = incl;
			__syncthreads();
			if (tid < RADIX) {
				u32 off = 0;
#pragma unroll
				for (int w = 0; w < RADIX / 32; ++w) if (w < warp) off += s_misc[1 + w];
				const u32 ds = off + incl - count;
				s_dstart[tid] = ds;
				u32 c[WARPS];
#pragma unroll
				for (int w = 0; w < WARPS; ++w) c[w] = whist[w * RADIX + tid];
				u32 run = ds;
#pragma unroll
				for (int w = 0; w < WARPS; ++w) { whist[w * RADIX + tid] = run; run += c[w]; }
			}
		};
		/* ---- stage the tile in digit order.  VERIFY_KEYS (keys only, identity key): the keys
		 *      themselves prove a correct pass, nothing else is staged.  Otherwise the
		 *      (digit, index-in-tile) word is staged beside the key. */
		auto stage = [&]() {
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 local = wbase + i * 32u;
				if (FULL || local < tile_count) {
					const u32 d = digit_of(key[i], local);
					const u32 p = wh[d] + ((pos2[i >> 1] >> (16 * (i & 1))) & 0xffffu);
					skeys[p] = key[i];
					if (USE_INFO) sinfo[p] = (d << 16) | local;
					if (HAS_VAL) svals[p] = val[i];
				}
			}
		};
		/* ---- write out: element j of the staged tile -> goff[digit] + j.  Returns whether the
		 *      staged order was seen to violate the pass invariant:
		 *      with the info word: (digit, index-in-tile) strictly increasing == stable;
		 *      keys only: the input of pass p is sorted on the bits below start_bit, so the
		 *      output is correct iff (key & low_mask) is non-decreasing along the staged tile,
		 *      low_mask covering this digit and everything below it. */
		auto write_out = [&](bool verify) -> bool {
			bool bad = false;
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 j = (u32) tid + i * THREADS;
				if (FULL || j < tile_count) {
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
					const LbT o = s_goff[d] + (LbT) j;
					out[o] = k;
					if (HAS_VAL) vout[o] = svals[j];
				}
			}
			return bad;
		};

		if (prof_on) { if (tid == 0 && key[IPT - 1] == ElemT(0x5a5a5a5a)) prof[15] = 1; mark(1); }  /* loads landed */
		if (RANK_MODE == RANK_ATOMIC) rank_atomic(); else rank_ballot();
		__syncthreads();
		mark(2);

		const u32 count = digit_count();
		/* publish the tile aggregate as early as possible */
		LbT* lb_mine = lookback + (size_t) tile * RADIX;
		if (tid < RADIX) {
			if (tile == 0) st_relaxed(lb_mine + tid, (LbT) (Lb<LbT>::PREFIX | (LbT) count));
			else st_relaxed(lb_mine + tid, (LbT) (Lb<LbT>::AGG | (LbT) count));
		}
		digit_starts(count);
		__syncthreads();
		mark(3);

		/* Staging needs tile-local offsets only.  The warps that own a digit (tid < RADIX)
		 * resolve their look-back first -- so the tile's inclusive prefix is published as
		 * early as possible, which keeps every successor's walk short -- while the other
		 * warps already stage their keys. */
		/* ---- decoupled look-back, one digit per thread */
		auto look_back = [&]() {
			LbT excl = 0;
			if (tile > 0) {
				long long p = (long long) tile - 1;
				unsigned spins = 0;
				bool done = false;
				/* A window of predecessor words is loaded at once (all loads in flight
				 * together) and consumed nearest first.  While tiles wait here, their
				 * successors must walk over them, so a short window makes the backlog --
				 * and with it every walk -- longer; after a first small window the walk
				 * continues with wide ones. */
				auto walk = [&](auto batch_tag) {
					constexpr int B = decltype(batch_tag)::value;
					LbT w[B];
#pragma unroll
					for (int k = 0; k < B; ++k)
						w[k] = (p - k >= 0) ? ld_relaxed(lookback + (size_t) (p - k) * RADIX + tid) : (LbT) Lb<LbT>::PREFIX;
#pragma unroll
					for (int k = 0; k < B; ++k) {
						if (!done) {
							const LbT f = w[k] & Lb<LbT>::FLAGS;
							if (f == 0) {            /* not published yet: retry from here */
								if (++spins > SPIN_LIMIT) { atomicExch(err_flag, 1); done = true; }
								break;
							}
							excl += w[k] & Lb<LbT>::VAL;
							--p;
							if (f == Lb<LbT>::PREFIX) done = true;
						}
					}
				};
				const long long lb_t0 = prof_on ? clock64() : 0;
				u32 rounds = 1;
				walk(std::integral_constant<int, LBF>{});
				while (!done) { walk(std::integral_constant<int, LBN>{}); ++rounds; }
				if (prof_on && tid == 0) {
					atomicAdd(prof + 7, (u64) (clock64() - lb_t0));   /* cycles inside the walk */
					atomicAdd(prof + 8, (u64) rounds);                  /* window loads */
					atomicAdd(prof + 9, (u64) spins);                   /* unpublished words met */
					atomicAdd(prof + 10, (u64) ((long long) tile - 1 - p));  /* predecessors consumed */
				}
				st_relaxed(lb_mine + tid, (LbT) (Lb<LbT>::PREFIX | ((excl + count) & Lb<LbT>::VAL)));
			}
			/* modular arithmetic in LbT: the final index goff[d] + j is < n */
			s_goff[tid] = (LbT) bins_base[tid] + excl - (LbT) s_dstart[tid];
		};
		if (tid < RADIX) look_back();
		stage();
		mark(4);
		__syncthreads();
		mark(5);

		const bool bad = write_out(VERIFY);
		mark(6);

		if (VERIFY) {
			if (__syncthreads_or(bad ? 1 : 0)) {
				/* The atomic ranks were not in lane order somewhere in this tile: redo the tile
				 * with the ballot ranks.  Digit counts, hence every offset, are unchanged. */
				if (tid == 0) atomicAdd(err_flag + 1, 1);
				for (int i = tid; i < WARPS * RADIX; i += THREADS) whist[i] = 0;
				__syncthreads();
#pragma unroll
				for (int i = 0; i < (IPT + 1) / 2; ++i) pos2[i] = 0;
				/* keys (and payloads) are re-read: their registers were released after staging */
#pragma unroll
				for (int i = 0; i < IPT; ++i) {
					const u32 local = wbase + i * 32u;
					if (FULL || local < tile_count) {
						key[i] = in[tile_base + local];
						if (HAS_VAL) val[i] = vin[tile_base + local];
					}
				}
				rank_ballot();
				__syncthreads();
				const u32 count2 = digit_count();
				digit_starts(count2);
				__syncthreads();
				stage();
				__syncthreads();
				write_out(false);
			}
		}
	};
	if (full) body(std::true_type{}); else body(std::false_type{});
}

#include "radix_pp.cuh"

/* ------------------------------------------------------------- host side */

template <typename ElemT, bool HAS_VAL> struct TileCfg {
	/* keys-only: 512 x 16 (4 B and narrower), 512 x 8 (8 B); with payload: 512 x 12 / 512 x 8 */
	static const int THREADS = 512;
	static const int IPT = HAS_VAL ? (sizeof(ElemT) == 8 ? 6 : 8) : (sizeof(ElemT) == 8 ? 10 : 16);
};

template <typename ElemT, bool HAS_VAL, int THREADS, int IPT, bool USE_INFO = true>
constexpr size_t onesweep_smem() {
	return (size_t) (THREADS / 32) * RADIX * 4 + RADIX * 4 + RADIX * 8 + 16 * 4 +
		(size_t) THREADS * IPT * sizeof(ElemT) + (USE_INFO || HAS_VAL ? (size_t) THREADS * IPT * 4 : 0) +
		(HAS_VAL ? (size_t) THREADS * IPT * 4 : 0);
}

} // namespace

struct CloRadixState {
	CloScratch aux_keys;     /* ping-pong partner of the key buffer */
	CloScratch aux_vals;
	CloScratch work;         /* [err | ghist | bins_base | tickets | lookback...] */
	CloScratch pp;           /* AGG + PREF words of the persistent kernel (self-cleaning) */
	int kernel_pp = 1;       /* CLO_RADIX_KERNEL=classic selects the one-tile-per-CTA kernel */
	int kernel_v6 = 1;       /* keys-only sorts use the two-barrier kernel; CLO_RADIX_KERNEL=pp|classic turn it off */
	int force_wide = 0;      /* CLO_RADIX_WIDE=1: 64-bit look-back words whatever n is (test hook for the >= 2^31 path) */
	int rank_atomic = 1;     /* CLO_RADIX_RANK=ballot selects the ballot ranks */
	int cfg = 0;
	/* optional per-kernel timing of the last call (clo_radix_set_timing) */
	/* sticky status of the last call: the device flag is copied to pinned host memory behind every
	 * call (no synchronisation); the next call looks at it once that copy has completed */
	int* h_status = nullptr;
	cudaEvent_t status_evt = nullptr;
	bool status_pending = false;
	int timing = 0;
	int n_marks = 0;
	cudaEvent_t marks[2 + MAX_PASSES + 2] = {};
	void mark(cudaStream_t stream) {
		if (!timing || n_marks >= (int) (sizeof(marks) / sizeof(marks[0]))) return;
		if (!marks[n_marks]) cudaEventCreate(&marks[n_marks]);
		cudaEventRecord(marks[n_marks++], stream);
	}
};

CloRadixState* clo_radix_state_new() {
	CloRadixState* st = new CloRadixState();
	const char* e = getenv("CLO_RADIX_RANK");
	st->rank_atomic = (e && strcmp(e, "ballot") == 0) ? 0 : 1;
	const char* kk = getenv("CLO_RADIX_KERNEL");
	st->kernel_pp = (kk && strcmp(kk, "classic") == 0) ? 0 : 1;
	st->kernel_v6 = (kk && (strcmp(kk, "classic") == 0 || strcmp(kk, "pp") == 0)) ? 0 : 1;
	const char* ppf = getenv("CLO_RADIX_PP_FLAGS");
	g_pp_flags = (ppf && *ppf) ? atoi(ppf) : 0;
	const char* pf = getenv("CLO_RADIX_PROFILE");
	g_radix_profile = (pf && *pf == '1') ? 1 : 0;
	const char* fw = getenv("CLO_RADIX_WIDE");
	st->force_wide = (fw && *fw == '1') ? 1 : 0;
	const char* c = getenv("CLO_RADIX_CFG");
	st->cfg = (c && *c) ? atoi(c) : 0;
	return st;
}

void clo_radix_set_timing(CloRadixState* st, int on) { if (st) st->timing = on; }

/* durations (ms) of the last call: out[0] = histogram + bin scan, out[1..passes] = the
 * onesweep passes; returns the number of values, 0 when timing was off */
int clo_radix_get_timing(CloRadixState* st, float* out, int cap) {
	if (!st || st->n_marks < 2) return 0;
	if (cudaEventSynchronize(st->marks[st->n_marks - 1]) != cudaSuccess) return 0;
	int k = 0;
	for (int i = 0; i + 1 < st->n_marks && k < cap; ++i, ++k)
		if (cudaEventElapsedTime(&out[k], st->marks[i], st->marks[i + 1]) != cudaSuccess) return 0;
	return k;
}

void clo_radix_state_free(CloRadixState* st) {
	if (!st) return;
	if (st->h_status) cudaFreeHost(st->h_status);
	if (st->status_evt) cudaEventDestroy(st->status_evt);
	for (cudaEvent_t e : st->marks) if (e) cudaEventDestroy(e);
	st->aux_keys.release(); st->aux_vals.release(); st->work.release(); st->pp.release();
	delete st;
}

int clo_radix_debug(CloRadixState* st, cudaStream_t stream, unsigned long long out[18]) {
	for (int i = 0; i < 18; ++i) out[i] = 0;
	if (!st || !st->work.ptr) return 0;
	int hdr[64];
	if (cudaMemcpyAsync(hdr, st->work.ptr, sizeof(hdr), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -1;
	if (cudaStreamSynchronize(stream) != cudaSuccess) return -1;
	out[0] = (unsigned long long) hdr[0];
	out[1] = (unsigned long long) hdr[1];
	memcpy(out + 2, hdr + 16, 16 * sizeof(unsigned long long));
	return 0;
}

int clo_radix_status(CloRadixState* st, cudaStream_t stream) {
	if (!st || !st->work.ptr) return 0;
	int flag = 0;
	if (cudaMemcpyAsync(&flag, st->work.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -1;
	if (cudaStreamSynchronize(stream) != cudaSuccess) return -1;
	if (flag != 0) {
		/* reported now: do not report it again on the next call, and start from clean words */
		if (st->pp.ptr) cudaMemsetAsync(st->pp.ptr, 0, st->pp.size, stream);
		cudaMemsetAsync(st->work.ptr, 0, sizeof(int), stream);
		if (st->h_status) *st->h_status = 0;
		st->status_pending = false;
	}
	return flag;
}

namespace {

const size_t WORK_HDR = 256;                              /* err flag lives here */
const size_t GHIST_BYTES = MAX_PASSES * RADIX * sizeof(u64);
const size_t TICKET_BYTES = 256;

struct WorkLayout {
	int* err; u64* ghist; u64* bins; u32* tickets; void* lookback; size_t zero_bytes;
};

cudaError_t prepare_work(CloRadixState* st, size_t tiles, int passes, size_t lb_word, WorkLayout& L, cudaStream_t stream) {
	const size_t lb_bytes = tiles * RADIX * lb_word * (size_t) passes;
	const size_t total = WORK_HDR + GHIST_BYTES + GHIST_BYTES + TICKET_BYTES + lb_bytes;
	cudaError_t e = st->work.reserve(total + total / 8);
	if (e != cudaSuccess) return e;
	char* base = (char*) st->work.ptr;
	L.err = (int*) base;
	L.ghist = (u64*) (base + WORK_HDR);
	L.tickets = (u32*) (base + WORK_HDR + GHIST_BYTES);
	L.lookback = base + WORK_HDR + GHIST_BYTES + TICKET_BYTES;
	L.bins = (u64*) ((char*) L.lookback + lb_bytes);
	/* everything up to the end of the look-back words is zeroed each call */
	L.zero_bytes = WORK_HDR + GHIST_BYTES + TICKET_BYTES + lb_bytes;
	return cudaMemsetAsync(base, 0, L.zero_bytes, stream);
}

template <typename ElemT, bool HAS_VAL, bool IDENTITY, typename LbT, int RANK_MODE, int THREADS, int IPT, int LBF = LB_FIRST, int LBN = LB_NEXT>
cudaError_t launch_onesweep(const ElemT* in, ElemT* out, const u32* vin, u32* vout, size_t n,
		LbT* lookback, u32* ticket, const u64* bins, u32 start_bit, u32 dmask, const CloKeySpec& ks,
		int* err, cudaStream_t stream) {
	constexpr size_t SMEM = onesweep_smem<ElemT, HAS_VAL, THREADS, IPT, (HAS_VAL || !IDENTITY)>();
	auto kern = clo_radix_onesweep<ElemT, HAS_VAL, IDENTITY, false, LbT, THREADS, IPT, RANK_MODE, LBF, LBN>;
	static bool configured[64] = {};
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || !configured[dev]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) configured[dev] = true;
	}
	const size_t tiles = (n + (size_t) THREADS * IPT - 1) / ((size_t) THREADS * IPT);
	SplitterArgs sp = { nullptr, nullptr, 0, 0 };
	kern<<<(unsigned) tiles, THREADS, SMEM, stream>>>(in, out, vin, vout, n, lookback, ticket, bins,
		start_bit, dmask, ks, sp, err, g_radix_profile);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

template <typename ElemT, bool HAS_VAL, bool IDENTITY, typename LbT, int RANK_MODE, int THREADS, int IPT>
cudaError_t launch_onesweep_pp(const ElemT* in, ElemT* out, const u32* vin, u32* vout, size_t n,
		LbT* agg, LbT* pref, u32* ticket, const u64* bins, u32 start_bit, u32 dmask, const CloKeySpec& ks,
		int* err, int sm_count, cudaStream_t stream) {
	constexpr size_t SMEM = onesweep_pp_smem<ElemT, HAS_VAL, IDENTITY, THREADS, IPT, LbT>();
	auto kern = clo_radix_onesweep_pp<ElemT, HAS_VAL, IDENTITY, LbT, THREADS, IPT, RANK_MODE>;
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
	const size_t tiles = (n + (size_t) THREADS * IPT - 1) / ((size_t) THREADS * IPT);
	size_t workers = (size_t) sm_count * ctas_per_sm[dev];
	workers = workers > (size_t) PP_NUM_PROP ? workers - PP_NUM_PROP : 1;
	if (workers > tiles) workers = tiles;
	kern<<<(unsigned) (PP_NUM_PROP + workers), THREADS, SMEM, stream>>>(in, out, vin, vout, n, (u32) tiles,
		agg, pref, ticket, bins, start_bit, dmask, ks, err, g_radix_profile, g_pp_flags);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

template <typename ElemT, bool IDENTITY, int PASSES, int COLS>
cudaError_t launch_histogram_p(const ElemT* src, size_t n, u64* ghist, const PassCfg& cfg, const CloKeySpec& ks,
		int vec_ok, int sm_count, cudaStream_t stream) {
	constexpr int THREADS = 1024;
	constexpr size_t SMEM = (size_t) PASSES * RADIX * COLS * sizeof(u32);
	auto kern = clo_radix_histogram<ElemT, IDENTITY, PASSES, COLS, THREADS>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM);
	if (e != cudaSuccess) return e;
	size_t want = (n + (size_t) THREADS * 16 - 1) / ((size_t) THREADS * 16);
	const unsigned blocks = (unsigned) (want < (size_t) sm_count ? (want ? want : 1) : (size_t) sm_count);
	kern<<<blocks, THREADS, SMEM, stream>>>(src, n, ghist, cfg, ks, vec_ok);
	return cudaGetLastError();
}

template <typename ElemT, bool IDENTITY>
cudaError_t launch_histogram(const ElemT* src, size_t n, u64* ghist, const PassCfg& cfg, const CloKeySpec& ks,
		int vec_ok, int sm_count, cudaStream_t stream) {
	/* 32 columns (conflict free) up to 4 passes = 128 KB; 16 columns for 5..8 passes */
	switch (cfg.passes) {
	case 1: return launch_histogram_p<ElemT, IDENTITY, 1, 32>(src, n, ghist, cfg, ks, vec_ok, sm_count, stream);
	case 2: return launch_histogram_p<ElemT, IDENTITY, 2, 32>(src, n, ghist, cfg, ks, vec_ok, sm_count, stream);
	case 3: case 4: return launch_histogram_p<ElemT, IDENTITY, 4, 32>(src, n, ghist, cfg, ks, vec_ok, sm_count, stream);
	default: return launch_histogram_p<ElemT, IDENTITY, 8, 16>(src, n, ghist, cfg, ks, vec_ok, sm_count, stream);
	}
}

template <typename ElemT, bool HAS_VAL, bool IDENTITY, int THREADS, int IPT, int LBF = LB_FIRST, int LBN = LB_NEXT>
cudaError_t radix_sort_cfg(CloRadixState* st, int sm_count, const CloKeySpec& ks, u32 sorted_bits,
		const ElemT* src, ElemT* dst, const u32* vsrc, u32* vdst, size_t n, cudaStream_t stream) {
	constexpr size_t TILE = (size_t) THREADS * IPT;
	cudaError_t e;
	PassCfg cfg;
	cfg.passes = (int) ((sorted_bits + RADIX_BITS - 1) / RADIX_BITS);
	for (int p = 0; p < MAX_PASSES; ++p) { cfg.start_bit[p] = 0; cfg.dmask[p] = 0; }
	for (int p = 0; p < cfg.passes; ++p) {
		const u32 rem = sorted_bits - (u32) p * RADIX_BITS;
		cfg.start_bit[p] = (u32) p * RADIX_BITS;
		cfg.dmask[p] = (1u << (rem < (u32) RADIX_BITS ? rem : (u32) RADIX_BITS)) - 1;
	}
	const size_t tiles = (n + TILE - 1) / TILE;
	const bool use_pp = st->kernel_pp != 0;
	/* look-back words: the classic kernel keeps 30 value bits, the persistent ones 31 (every
	 * prefix and offset is a count of keys, i.e. < n) */
	const bool wide = st->force_wide || n >= (use_pp ? (1ull << 31) : (1ull << 30));
	WorkLayout L;
	if ((e = prepare_work(st, use_pp ? 0 : tiles, cfg.passes, wide ? 8 : 4, L, stream)) != cudaSuccess) return e;
	if (use_pp) {
		/* AGG + PREF words, shared by all passes; zeroed when (re)allocated, self-cleaning after */
		const size_t need = 2 * tiles * RADIX * (wide ? 8 : 4);
		if (need > st->pp.size) {
			if ((e = st->pp.reserve(need + need / 4)) != cudaSuccess) return e;
			if ((e = cudaMemsetAsync(st->pp.ptr, 0, st->pp.size, stream)) != cudaSuccess) return e;
		}
	}
	if ((e = st->aux_keys.reserve(n * sizeof(ElemT))) != cudaSuccess) return e;
	if (HAS_VAL && (e = st->aux_vals.reserve(n * sizeof(u32))) != cudaSuccess) return e;
	ElemT* aux = (ElemT*) st->aux_keys.ptr;
	u32* vaux = (u32*) st->aux_vals.ptr;

	st->n_marks = 0;
	st->mark(stream);
	/* histogram of every digit in one read of the keys: one persistent CTA per SM */
	{
		const int vec_ok = (reinterpret_cast<uintptr_t>(src) % 16) == 0;
		if ((e = launch_histogram<ElemT, IDENTITY>(src, n, L.ghist, cfg, ks, vec_ok, sm_count, stream)) != cudaSuccess) return e;
		clo_radix_scan_bins<<<cfg.passes, RADIX, 0, stream>>>(L.ghist, L.bins);
		CLO_COUNT_LAUNCH(2);
		if ((e = cudaGetLastError()) != cudaSuccess) return e;
	}
	st->mark(stream);

	/* ping-pong chain that ends in dst without touching src (unless in place):
	 * odd number of passes: src->dst->aux->dst...; even: src->aux->dst->aux->dst */
	const ElemT* cur = src; const u32* vcur = vsrc;
	if ((const void*) src == (const void*) dst && (cfg.passes & 1)) {
		/* in place with an odd number of passes: start the chain from the aux copy */
		if ((e = cudaMemcpyAsync(aux, src, n * sizeof(ElemT), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
		cur = aux;
		if (HAS_VAL) {
			if ((e = cudaMemcpyAsync(vaux, vsrc, n * sizeof(u32), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
			vcur = vaux;
		}
	}
	for (int p = 0; p < cfg.passes; ++p) {
		const bool to_dst = ((cfg.passes - 1 - p) % 2) == 0;
		ElemT* nxt = to_dst ? dst : aux;
		u32* vnxt = to_dst ? vdst : vaux;
		u32* ticket = L.tickets + p;
		bool done_v6 = false;
		if constexpr (IDENTITY && sizeof(ElemT) >= 4 && THREADS > RADIX) {
			if (use_pp && st->kernel_v6) {
				char* agg = (char*) st->pp.ptr;
				char* pref = agg + tiles * RADIX * (wide ? 8 : 4);
				e = clo_radix_v6_pass((int) sizeof(ElemT), wide ? 1 : 0, THREADS * IPT, cur, nxt, HAS_VAL ? vcur : nullptr, HAS_VAL ? vnxt : nullptr, n, agg, pref, ticket,
					L.bins + p * RADIX, cfg.start_bit[p], cfg.dmask[p], L.err, sm_count, g_radix_profile, g_pp_flags, stream);
				done_v6 = true;
			}
		}
		if (done_v6) {
		} else if (use_pp) {
			const u64* bins = L.bins + p * RADIX;
			if (wide) {
				u64* agg = (u64*) st->pp.ptr; u64* pref = agg + tiles * RADIX;
				e = st->rank_atomic
					? launch_onesweep_pp<ElemT, HAS_VAL, IDENTITY, u64, 1, THREADS, IPT>(cur, nxt, vcur, vnxt, n, agg, pref, ticket, bins, cfg.start_bit[p], cfg.dmask[p], ks, L.err, sm_count, stream)
					: launch_onesweep_pp<ElemT, HAS_VAL, IDENTITY, u64, 0, THREADS, IPT>(cur, nxt, vcur, vnxt, n, agg, pref, ticket, bins, cfg.start_bit[p], cfg.dmask[p], ks, L.err, sm_count, stream);
			} else {
				u32* agg = (u32*) st->pp.ptr; u32* pref = agg + tiles * RADIX;
				e = st->rank_atomic
					? launch_onesweep_pp<ElemT, HAS_VAL, IDENTITY, u32, 1, THREADS, IPT>(cur, nxt, vcur, vnxt, n, agg, pref, ticket, bins, cfg.start_bit[p], cfg.dmask[p], ks, L.err, sm_count, stream)
					: launch_onesweep_pp<ElemT, HAS_VAL, IDENTITY, u32, 0, THREADS, IPT>(cur, nxt, vcur, vnxt, n, agg, pref, ticket, bins, cfg.start_bit[p], cfg.dmask[p], ks, L.err, sm_count, stream);
			}
		} else if (wide) {
			u64* lb = (u64*) L.lookback + (size_t) p * tiles * RADIX;
			e = launch_onesweep<ElemT, HAS_VAL, IDENTITY, u64, 0, THREADS, IPT, LBF, LBN>(cur, nxt, vcur, vnxt, n, lb, ticket, L.bins + p * RADIX, cfg.start_bit[p], cfg.dmask[p], ks, L.err, stream);
		} else {
			u32* lb = (u32*) L.lookback + (size_t) p * tiles * RADIX;
			e = launch_onesweep<ElemT, HAS_VAL, IDENTITY, u32, 0, THREADS, IPT, LBF, LBN>(cur, nxt, vcur, vnxt, n, lb, ticket, L.bins + p * RADIX, cfg.start_bit[p], cfg.dmask[p], ks, L.err, stream);
		}
		if (e != cudaSuccess) return e;
		st->mark(stream);
		cur = nxt; vcur = vnxt;
	}
	return cudaSuccess;
}

template <typename ElemT, bool HAS_VAL, bool IDENTITY>
cudaError_t radix_sort_typed(CloRadixState* st, int sm_count, const CloKeySpec& ks, u32 sorted_bits,
		const ElemT* src, ElemT* dst, const u32* vsrc, u32* vdst, size_t n, cudaStream_t stream) {
	return radix_sort_cfg<ElemT, HAS_VAL, IDENTITY, TileCfg<ElemT, HAS_VAL>::THREADS, TileCfg<ElemT, HAS_VAL>::IPT>(
		st, sm_count, ks, sorted_bits, src, dst, vsrc, vdst, n, stream);
}

/* tile-shape variant of the headline case (u32 keys only, identity key): CLO_RADIX_CFG=1 */
template <>
cudaError_t radix_sort_typed<u32, false, true>(CloRadixState* st, int sm_count, const CloKeySpec& ks, u32 sorted_bits,
		const u32* src, u32* dst, const u32* vsrc, u32* vdst, size_t n, cudaStream_t stream) {
	if (st->cfg == 1)
		return radix_sort_cfg<u32, false, true, 256, 16>(st, sm_count, ks, sorted_bits, src, dst, vsrc, vdst, n, stream);
	return radix_sort_cfg<u32, false, true, 512, CLO_IPT_U32>(st, sm_count, ks, sorted_bits, src, dst, vsrc, vdst, n, stream);
}

template <typename ElemT>
cudaError_t radix_sort_elem(CloRadixState* st, int sm_count, const CloKeySpec& ks, u32 sorted_bits,
		const void* src, void* dst, const u32* vsrc, u32* vdst, size_t n, cudaStream_t stream) {
	const bool has_val = vsrc != nullptr;
	if (has_val) {
		if (ks.identity) return radix_sort_typed<ElemT, true, true>(st, sm_count, ks, sorted_bits, (const ElemT*) src, (ElemT*) dst, vsrc, vdst, n, stream);
		return radix_sort_typed<ElemT, true, false>(st, sm_count, ks, sorted_bits, (const ElemT*) src, (ElemT*) dst, vsrc, vdst, n, stream);
	}
	if (ks.identity) return radix_sort_typed<ElemT, false, true>(st, sm_count, ks, sorted_bits, (const ElemT*) src, (ElemT*) dst, vsrc, vdst, n, stream);
	return radix_sort_typed<ElemT, false, false>(st, sm_count, ks, sorted_bits, (const ElemT*) src, (ElemT*) dst, vsrc, vdst, n, stream);
}

} // namespace

cudaError_t clo_radix_sort(CloRadixState* st, int sm_count, size_t elem_size, const CloKeySpec& ks,
		uint32_t sorted_bits, const void* src, void* dst, const uint32_t* payload_src, uint32_t* payload_dst,
		size_t n, cudaStream_t stream, const char** err_msg) {
	if (n == 0) return cudaSuccess;
	if (n >= (1ull << 40)) { if (err_msg) *err_msg = "radix sort: too many elements"; return cudaErrorInvalidValue; }
	if (sorted_bits == 0 || sorted_bits > 64) { if (err_msg) *err_msg = "radix sort: invalid number of key bits"; return cudaErrorInvalidValue; }
	if (payload_src && elem_size < 4) { if (err_msg) *err_msg = "radix sort: payload needs 4- or 8-byte keys"; return cudaErrorInvalidValue; }
	/* A pass that ran into a look-back / propagator timeout (a grid that was not fully resident)
	 * leaves wrong output and stale AGG / PREF words.  Callers of the device-data entry points never
	 * synchronise, so the flag of the PREVIOUS call is reported here: that call's output is invalid. */
	if (st->status_pending && cudaEventQuery(st->status_evt) == cudaSuccess) {
		st->status_pending = false;
		if (*st->h_status != 0) {
			*st->h_status = 0;
			if (st->pp.ptr) cudaMemsetAsync(st->pp.ptr, 0, st->pp.size, stream);
			if (err_msg) *err_msg = "radix sort: the previous sort on this sorter timed out waiting for tile prefixes; its output is invalid";
			return cudaErrorLaunchFailure;
		}
	}
	cudaError_t rc;
	switch (elem_size) {
	case 1: rc = radix_sort_elem<unsigned char>(st, sm_count, ks, sorted_bits, src, dst, nullptr, nullptr, n, stream); break;
	case 2: rc = radix_sort_elem<unsigned short>(st, sm_count, ks, sorted_bits, src, dst, nullptr, nullptr, n, stream); break;
	case 4: rc = radix_sort_elem<u32>(st, sm_count, ks, sorted_bits, src, dst, payload_src, payload_dst, n, stream); break;
	case 8: rc = radix_sort_elem<u64>(st, sm_count, ks, sorted_bits, src, dst, payload_src, payload_dst, n, stream); break;
	default: if (err_msg) *err_msg = "radix sort: unsupported element size"; return cudaErrorInvalidValue;
	}
	if (rc == cudaSuccess && st->work.ptr) {
		if (!st->h_status && cudaMallocHost((void**) &st->h_status, sizeof(int)) == cudaSuccess) *st->h_status = 0;
		if (!st->status_evt) cudaEventCreateWithFlags(&st->status_evt, cudaEventDisableTiming);
		if (st->h_status && st->status_evt && !st->status_pending &&
			cudaMemcpyAsync(st->h_status, st->work.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
			cudaEventRecord(st->status_evt, stream) == cudaSuccess)
			st->status_pending = true;
	}
	return rc;
}

/* ------------------------------------------------------------- partition */

namespace {

template <typename ElemT, bool HAS_VAL>
cudaError_t partition_typed(CloRadixState* st, const ElemT* in, const u32* vin, ElemT* out, u32* vout, size_t n,
		u64 gidx0, const void* sk, const u64* si, u32 nparts, u64* counts_out, cudaStream_t stream);

/* bucket histogram: counts per bucket (same bucket function as the onesweep variant) */
template <typename ElemT, int THREADS>
__global__ void __launch_bounds__(THREADS)
clo_partition_histogram(const ElemT* __restrict__ in, size_t n, SplitterArgs sp, u64* __restrict__ ghist) {
	__shared__ u32 sh[16];
	__shared__ ElemT s_key[15];
	__shared__ u64 s_idx[15];
	if (threadIdx.x < 16) sh[threadIdx.x] = 0;
	if (threadIdx.x < sp.count) {
		s_key[threadIdx.x] = reinterpret_cast<const ElemT*>(sp.keys)[threadIdx.x];
		s_idx[threadIdx.x] = sp.idx[threadIdx.x];
	}
	__syncthreads();
	u32 local[16];
#pragma unroll
	for (int b = 0; b < 16; ++b) local[b] = 0;
	for (size_t i = (size_t) blockIdx.x * THREADS + threadIdx.x; i < n; i += (size_t) gridDim.x * THREADS) {
		const ElemT k = __ldcs(in + i);
		const u64 g = sp.gidx0 + i;
		u32 b = 0;
		for (u32 s = 0; s < sp.count; ++s)
			if (s_key[s] < k || (s_key[s] == k && s_idx[s] <= g)) ++b;
#pragma unroll
		for (int q = 0; q < 16; ++q) if (q == (int) b) local[q]++;
	}
#pragma unroll
	for (int b = 0; b < 16; ++b) {
		const u32 t = warp_reduce_sum<u32>(local[b]);
		if ((threadIdx.x & 31) == 0 && t) atomicAdd(&sh[b], t);
	}
	__syncthreads();
	if (threadIdx.x < 16 && sh[threadIdx.x]) atomicAdd(&ghist[threadIdx.x], (u64) sh[threadIdx.x]);
}

/* bins_base[b] = exclusive scan of the bucket counts; counts_out[b] = count */
__global__ void clo_partition_scan_bins(const u64* __restrict__ ghist, u64* __restrict__ bins_base,
		u64* __restrict__ counts_out, u32 nparts) {
	if (threadIdx.x == 0 && blockIdx.x == 0) {
		u64 run = 0;
		for (u32 b = 0; b < (u32) RADIX; ++b) {
			const u64 c = b < 16 ? ghist[b] : 0;
			bins_base[b] = run;
			run += c;
			if (b < nparts && counts_out) counts_out[b] = c;
		}
	}
}

template <typename ElemT, bool HAS_VAL>
cudaError_t partition_typed(CloRadixState* st, const ElemT* in, const u32* vin, ElemT* out, u32* vout, size_t n,
		u64 gidx0, const void* sk, const u64* si, u32 nparts, u64* counts_out, cudaStream_t stream) {
	constexpr int THREADS = TileCfg<ElemT, HAS_VAL>::THREADS;
	constexpr int IPT = TileCfg<ElemT, HAS_VAL>::IPT;
	constexpr size_t TILE = (size_t) THREADS * IPT;
	constexpr size_t SMEM = onesweep_smem<ElemT, HAS_VAL, THREADS, IPT>();
	const size_t tiles = (n + TILE - 1) / TILE;
	const bool wide = n >= (1ull << 30);
	WorkLayout L;
	cudaError_t e;
	if ((e = prepare_work(st, tiles, 1, wide ? 8 : 4, L, stream)) != cudaSuccess) return e;
	SplitterArgs sp = { sk, si, nparts - 1, gidx0 };
	CloKeySpec ks = {};
	ks.identity = 1;
	const unsigned hb = (unsigned) (tiles < 1184 ? (tiles ? tiles : 1) : 1184);
	clo_partition_histogram<ElemT, 256><<<hb, 256, 0, stream>>>(in, n, sp, L.ghist);
	clo_partition_scan_bins<<<1, 32, 0, stream>>>(L.ghist, L.bins, counts_out, nparts);
	CLO_COUNT_LAUNCH(2);
	if ((e = cudaGetLastError()) != cudaSuccess) return e;
	if (n == 0) return cudaSuccess;
	if (wide) {
		auto kern = clo_radix_onesweep<ElemT, HAS_VAL, true, true, u64, THREADS, IPT, RANK_BALLOT>;
		if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM)) != cudaSuccess) return e;
		kern<<<(unsigned) tiles, THREADS, SMEM, stream>>>(in, out, vin, vout, n, (u64*) L.lookback, L.tickets, L.bins, 0, 0xff, ks, sp, L.err, 0);
	} else {
		auto kern = clo_radix_onesweep<ElemT, HAS_VAL, true, true, u32, THREADS, IPT, RANK_BALLOT>;
		if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM)) != cudaSuccess) return e;
		kern<<<(unsigned) tiles, THREADS, SMEM, stream>>>(in, out, vin, vout, n, (u32*) L.lookback, L.tickets, L.bins, 0, 0xff, ks, sp, L.err, 0);
	}
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

} // namespace

static int clo_current_sm_count() {
	int dev = 0, sms = 0;
	cudaGetDevice(&dev);
	if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
	return sms;
}

cudaError_t clo_radix_partition_count(CloRadixState* st, size_t elem_size, const void* keys_in, size_t n, uint64_t gidx0,
		const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out,
		cudaStream_t stream, const char** err_msg) {
	return clo_partition_count_stage(st->work, elem_size, keys_in, n, gidx0, splitter_keys, splitter_idx, nparts,
		counts_out, clo_current_sm_count(), stream, err_msg);
}

cudaError_t clo_radix_partition_scatter(CloRadixState* st, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		size_t n, uint64_t gidx0, const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts,
		const uint64_t* first_slot, void* const* dests, void* const* vdests, const int* ok,
		cudaStream_t stream, const char** err_msg) {
	return clo_partition_scatter_stage(st->work, elem_size, keys_in, payload_in, n, gidx0, splitter_keys, splitter_idx,
		nparts, first_slot, dests, vdests, ok, clo_current_sm_count(), stream, err_msg);
}

cudaError_t clo_radix_partition(CloRadixState* st, size_t elem_size, const void* keys_in,
		const uint32_t* payload_in, void* keys_out, uint32_t* payload_out, size_t n, uint64_t gidx0,
		const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out,
		cudaStream_t stream, const char** err_msg) {
	if (nparts < 1 || nparts > 16) { if (err_msg) *err_msg = "partition: nparts must be in [1,16]"; return cudaErrorInvalidValue; }
	if (n >= (1ull << 40)) { if (err_msg) *err_msg = "partition: too many elements"; return cudaErrorInvalidValue; }
	{
		/* default: the chunk-per-warp partition (partition.cu); CLO_PARTITION=onesweep keeps the
		 * splitter-as-digit onesweep variant below */
		const char* pk = getenv("CLO_PARTITION");
		if (!(pk && strcmp(pk, "onesweep") == 0) && n < (1ull << 32)) {
			int dev = 0, sms = 0;
			cudaGetDevice(&dev);
			if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
			return clo_partition_v2(st->work, elem_size, keys_in, payload_in, keys_out, payload_out, n, gidx0,
				splitter_keys, splitter_idx, nparts, counts_out, sms, stream, err_msg);
		}
	}
	const bool has_val = payload_in != nullptr;
	if (elem_size == 4) {
		if (has_val) return partition_typed<u32, true>(st, (const u32*) keys_in, payload_in, (u32*) keys_out, payload_out, n, (u64) gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, stream);
		return partition_typed<u32, false>(st, (const u32*) keys_in, nullptr, (u32*) keys_out, nullptr, n, (u64) gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, stream);
	}
	if (elem_size == 8) {
		if (has_val) return partition_typed<u64, true>(st, (const u64*) keys_in, payload_in, (u64*) keys_out, payload_out, n, (u64) gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, stream);
		return partition_typed<u64, false>(st, (const u64*) keys_in, nullptr, (u64*) keys_out, nullptr, n, (u64) gidx0, splitter_keys, (const u64*) splitter_idx, nparts, (u64*) counts_out, stream);
	}
	if (err_msg) *err_msg = "partition: keys must be 4 or 8 bytes";
	return cudaErrorInvalidValue;
}
