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
 *
 * Round 2 (ncu: the pass is bound by the shared-memory data pipe, 19 wavefronts per 32 keys):
 *   - the warp histograms are PACKED, two 16-bit digit counters per word (a staged tile has
 *     < 2^16 slots): 128 digit threads own two digits each, so the digit phase and the row
 *     clearing touch half the words;
 *   - VERIFY is a template flag.  The per-tile stability check costs one more LDS per key; the
 *     launcher runs clo_radix_atomic_order_selftest ONCE per device (the same instruction
 *     forms on heavily colliding addresses, checked against __match_any_sync ranks) and uses
 *     the unverified instance only when the device serves same-address lanes in lane order.
 *     CLO_RADIX_VERIFY=1 forces the verified instance; flag 16 plants a real inversion in
 *     front of the detector (tests).
 */
#ifndef CLO_NO_SWIZZLE
#define CLO_NO_SWIZZLE 0
#endif
#ifndef CLO_RADIX_V6_CUH
#define CLO_RADIX_V6_CUH

/* location of the keys (and payloads) between two passes of one sort */
struct V6Chain { const void* cur; const void* vcur; };

/* after the last pass: the result belongs in dst; copies only when data-dependent skipping left
 * it elsewhere (16 bytes per thread and step; buffers from cudaMalloc are always aligned) */
__global__ void __launch_bounds__(512)
clo_radix_chain_fixup(const V6Chain* __restrict__ chain, void* __restrict__ dst, void* __restrict__ vdst, size_t key_bytes, size_t val_bytes) {
	auto copy = [&](const void* src, void* to, size_t bytes) {
		if (src == (const void*) to || bytes == 0) return;
		const size_t w = (size_t) blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t) gridDim.x * blockDim.x;
		if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(to)) & 15) == 0) {
			const size_t nv = bytes / 16;
			for (size_t i = w; i < nv; i += stride)
				__stcs(reinterpret_cast<uint4*>(to) + i, __ldcs(reinterpret_cast<const uint4*>(src) + i));
			for (size_t i = nv * 16 + w; i < bytes; i += stride)
				reinterpret_cast<unsigned char*>(to)[i] = reinterpret_cast<const unsigned char*>(src)[i];
		} else {
			for (size_t i = w; i < bytes; i += stride)
				reinterpret_cast<unsigned char*>(to)[i] = reinterpret_cast<const unsigned char*>(src)[i];
		}
	};
	copy(chain->cur, dst, key_bytes);
	if (vdst) copy(chain->vcur, vdst, val_bytes);
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
	asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

template <typename ElemT>
__device__ __forceinline__ u32 v6_digit(ElemT k, u32 start_bit, u32 dmask) {
	if (sizeof(ElemT) <= 4) return (((u32) k) >> start_bit) & dmask;
	return ((u32) (((u64) k) >> start_bit)) & dmask;
}

/* Rare path, kept out of line so that it costs the hot loop no registers.  Tile t (already
 * written out in a wrong intra-digit order) is ranked again with ballots and scattered
 * straight to its final positions.  scratch is the tile's own staging buffer ([WARPS][RADIX]
 * words are used); goff and ds are still those of tile t.  Called by all threads of the CTA. */
template <typename ElemT, typename LbT, int THREADS, int IPT, bool HAS_VAL>
__device__ __noinline__ void v6_repair(const ElemT* __restrict__ in, ElemT* __restrict__ out,
		const u32* __restrict__ vin, u32* __restrict__ vout, size_t n, u32 t,
		u32* scratch, const LbT* goff, const u32* ds, u32 start_bit, u32 dmask, int* err_flag) {
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * IPT;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const size_t base = (size_t) t * TILE;
	const u32 cnt = (base + TILE <= n) ? (u32) TILE : (u32) (n - base);
	const u32 wbase = (u32) warp * 32u * IPT + lane;
	if (tid == 0) atomicAdd(err_flag + 1, 1);
	for (int i = tid; i < WARPS * RADIX; i += THREADS) scratch[i] = 0;
	__syncthreads();
	u32* sw = scratch + warp * RADIX;
	ElemT key[IPT];
	u32 val[HAS_VAL ? IPT : 1];
	u32 rank[IPT];
#pragma unroll 1
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
#pragma unroll 1
	for (int i = 0; i < IPT; ++i) {
		const u32 local = wbase + i * 32u;
		if (local < cnt) {
			const u32 d = v6_digit<ElemT>(key[i], start_bit, dmask);
			const LbT o = goff[d] + (LbT) (ds[d] + sw[d] + rank[i]);
			out[o] = key[i];
			if (HAS_VAL) vout[o] = vin[base + local];
		}
	}
	__syncthreads();
}

/* propagator groups (32 digits each) per CTA and propagator CTAs, by CTA size */
__host__ __device__ constexpr int v6_prop_groups(int threads, int word_bytes) { return (threads >= 512 && word_bytes == 4) ? 2 : 1; }
__host__ __device__ constexpr int v6_num_prop(int threads, int word_bytes) { return RADIX / 32 / v6_prop_groups(threads, word_bytes); }

template <typename ElemT, typename LbT, int THREADS, int IPT, bool HAS_VAL = false, bool VERIFY = true>
__global__ void __launch_bounds__(THREADS, (THREADS <= 384 ? 3 : 2))
clo_radix_onesweep_v6(const ElemT* in_arg, ElemT* out_arg,
		const u32* vin_arg, u32* vout_arg, size_t n, u32 num_tiles,
		LbT* __restrict__ agg, LbT* __restrict__ pref, u32* __restrict__ ticket,
		const u64* __restrict__ bins_base, u32 start_bit, u32 dmask,
		int* __restrict__ err_flag, int prof_on, int flags,
		const V6Chain* __restrict__ chain_in, V6Chain* __restrict__ chain_out, ElemT* out_alt, u32* vout_alt) {
	/* Where the keys are.  A pass whose digit is the same for every key moves nothing, so which
	 * buffer holds the keys after p passes depends on the data: the launcher gives every pass
	 * the output it would use if all passes were real (out_arg) and the other buffer (out_alt);
	 * the pass reads the current location from the chain word its predecessor wrote, sorts into
	 * whichever of the two it is not reading, and tells its successor (block 0).  chain_in ==
	 * NULL: the first pass, or a caller that still moves the keys itself (identity = copy). */
	/* For 8-byte keys and for key + payload sorts the four resolved pointers live in SHARED memory,
	 * not in registers: kernel parameters cost no registers (constant bank), values loaded from the
	 * chain word do -- eight of them, which at the 64-register cap of this kernel meant more spills
	 * (pairs u64 + u32: 8.23 -> 7.69 ms per 2^27 with the pointers in shared memory; every use is per
	 * tile, so a broadcast load each time is free).  The 4-byte keys-only instance is the other way
	 * round (2.80 vs 2.92 ms per 2^28): it keeps them in registers. */
	constexpr bool PTRS_IN_SMEM = HAS_VAL || sizeof(ElemT) == 8;
	__shared__ struct { const ElemT* in; ElemT* out; const u32* vin; u32* vout; } s_p;
	const ElemT* r_in = nullptr; const u32* r_vin = nullptr; ElemT* r_out = nullptr; u32* r_vout = nullptr;
	if (PTRS_IN_SMEM) {
		if (threadIdx.x == 0) {
			const ElemT* in0 = chain_in ? (const ElemT*) chain_in->cur : in_arg;
			const bool swap = (const void*) out_arg == (const void*) in0 && out_alt;
			s_p.in = in0;
			s_p.vin = chain_in ? (const u32*) chain_in->vcur : vin_arg;
			s_p.out = swap ? out_alt : out_arg;
			s_p.vout = swap ? vout_alt : vout_arg;
		}
		__syncthreads();
	} else {
		r_in = chain_in ? (const ElemT*) chain_in->cur : in_arg;
		r_vin = chain_in ? (const u32*) chain_in->vcur : vin_arg;
		const bool swap = (const void*) out_arg == (const void*) r_in && out_alt;
		r_out = swap ? out_alt : out_arg;
		r_vout = swap ? vout_alt : vout_arg;
	}
	const ElemT* const& in = PTRS_IN_SMEM ? s_p.in : r_in;
	const u32* const& vin = PTRS_IN_SMEM ? s_p.vin : r_vin;
	ElemT* const& out = PTRS_IN_SMEM ? s_p.out : r_out;
	u32* const& vout = PTRS_IN_SMEM ? s_p.vout : r_vout;
	constexpr int WARPS = THREADS / 32;
	constexpr int TILE = THREADS * IPT;
	constexpr u32 NONE = 0xffffffffu;
	constexpr int DT = RADIX / 2;                  /* digit threads: two digits (one packed word) each */
	constexpr int ROWW = RADIX / 2;                /* packed words per warp row */
	constexpr int PREF_T = (THREADS - DT) >= RADIX ? RADIX : (THREADS - DT);   /* prefix threads: tid in [DT, DT + PREF_T) */
	static_assert(THREADS > DT && RADIX % PREF_T == 0, "128 digit threads + prefix threads owning whole digits");
	constexpr int DPT = RADIX / PREF_T;            /* digits per prefix thread */
	static_assert(ROWW == 128, "a lane clears its share of a row with one 16-byte store");
	static_assert((size_t) TILE * sizeof(ElemT) >= (size_t) WARPS * RADIX * 4, "repair scratch lives in a staging buffer");
	static_assert(TILE < 65536, "slots and indices in a tile must fit 16 bits");

	extern __shared__ __align__(16) unsigned char smem_raw[];

	/* A pass whose digit is the same for every key is the identity permutation (keys that are
	 * small, already grouped, or all equal): the global histogram says so up front, and the
	 * whole grid then just copies its tiles -- no ranking, no prefix traffic. */
	{
		__shared__ int s_trivial;
		if (threadIdx.x == 0) s_trivial = 0;
		__syncthreads();
		if (threadIdx.x < RADIX) {
			const u64 hi = threadIdx.x + 1 < RADIX ? bins_base[threadIdx.x + 1] : (u64) n;
			if (hi - bins_base[threadIdx.x] == (u64) n) s_trivial = 1;
		}
		__syncthreads();
		if (chain_out && blockIdx.x == 0 && threadIdx.x == 0) {
			chain_out->cur = s_trivial ? (const void*) in : (const void*) out;
			chain_out->vcur = s_trivial ? (const void*) vin : (const void*) vout;
		}
		if (s_trivial && chain_out) return;            /* nothing moves: the successor reads where we read */
		if (s_trivial) {
			constexpr int NP = v6_num_prop(THREADS, (int) sizeof(LbT));
			if (blockIdx.x < NP) return;
			/* 16 bytes per thread and step when the buffers allow it (cudaMalloc'ed ones always do) */
			auto copy = [&](const void* src, void* dst, size_t bytes) {
				const size_t w = (size_t) (blockIdx.x - NP) * THREADS + threadIdx.x, stride = (size_t) (gridDim.x - NP) * THREADS;
				if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
					const size_t nv = bytes / 16;
					for (size_t i = w; i < nv; i += stride)
						__stcs(reinterpret_cast<uint4*>(dst) + i, __ldcs(reinterpret_cast<const uint4*>(src) + i));
					for (size_t i = nv * 16 + w; i < bytes; i += stride)
						reinterpret_cast<unsigned char*>(dst)[i] = reinterpret_cast<const unsigned char*>(src)[i];
				} else {
					for (size_t i = w; i < bytes / 4; i += stride)
						reinterpret_cast<u32*>(dst)[i] = __ldcs(reinterpret_cast<const u32*>(src) + i);
				}
			};
			copy(in, out, n * sizeof(ElemT));
			if (HAS_VAL) copy(vin, vout, n * 4);
			return;
		}
	}

	/* The propagators are latency critical; a worker on the same SM puts its shared-memory
	 * traffic in front of every propagator load.  Each propagator publishes the id of its SM
	 * and a worker that finds itself on one of them leaves (tiles go by ticket: nothing is lost). */
	u32* prop_sm = reinterpret_cast<u32*>(err_flag + 48);       /* [PP_NUM_PROP], zeroed per call */
	u32 my_sm;
	asm volatile("mov.u32 %0, %%smid;" : "=r"(my_sm));
	constexpr int V6_PROP_GROUPS = v6_prop_groups(THREADS, (int) sizeof(LbT));
	constexpr int V6_NUM_PROP = v6_num_prop(THREADS, (int) sizeof(LbT));
	if (blockIdx.x < V6_NUM_PROP) {
		if (threadIdx.x == 0) st_relaxed(prop_sm + blockIdx.x, my_sm + 1u);
		pp_propagate2<LbT, THREADS, V6_PROP_GROUPS>(agg, pref, num_tiles, err_flag, smem_raw, prof_on, (int) blockIdx.x);
		return;
	}
	if (flags & 4) {
		__shared__ int s_leave;
		if (threadIdx.x == 0) {
			int leave = 0;
			for (int k = 0; k < V6_NUM_PROP; ++k) {
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

	u32* whist = reinterpret_cast<u32*>(smem_raw);                          /* [WARPS][ROWW] packed: digit d = half (d & 1) of word d >> 1 */
	u32* s_ds = whist + WARPS * ROWW;                                       /* [3][RADIX] digit starts (+1 pad) */
	u64* s_goff_raw = reinterpret_cast<u64*>(s_ds + 4 * RADIX);             /* [2][RADIX] u64-sized slots (s_ds has 3 live slots + 1 pad) */
	u32* s_misc = reinterpret_cast<u32*>(s_goff_raw + 2 * RADIX);           /* [32]: 0..7 scan, 8..9 tickets, 12..15 warp maxima, 16 heaviest digit */
	/* staging, per buffer: keys [TILE]; with a payload also values [TILE] and the index in tile
	 * of every staged element [TILE] (u16), which is what the stability check compares */
	constexpr size_t BUF_BYTES = (size_t) TILE * sizeof(ElemT) + (HAS_VAL ? (size_t) TILE * 6 : 0);
	unsigned char* s_buf_raw = reinterpret_cast<unsigned char*>(s_misc + 32);   /* [2][BUF_BYTES] */
	auto buf_keys = [&](int b) { return reinterpret_cast<ElemT*>(s_buf_raw + (size_t) b * BUF_BYTES); };
	auto buf_vals = [&](int b) { return reinterpret_cast<u32*>(s_buf_raw + (size_t) b * BUF_BYTES + (size_t) TILE * sizeof(ElemT)); };
	auto buf_info = [&](int b) { return reinterpret_cast<unsigned short*>(s_buf_raw + (size_t) b * BUF_BYTES + (size_t) TILE * (sizeof(ElemT) + 4)); };

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const ElemT low_mask = (ElemT) ((((ElemT) dmask) << start_bit) | ((((ElemT) 1) << start_bit) - 1));
	const u32 wbase = (u32) warp * 32u * IPT + lane;
	u32* wh = whist + warp * ROWW;
	/* the prefix threads own DPT digits each */
	const bool is_pref_thread = tid >= DT && tid < DT + PREF_T;
	LbT my_base[DPT];
#pragma unroll
	for (int j = 0; j < DPT; ++j) my_base[j] = is_pref_thread ? (LbT) bins_base[tid - DT + j * PREF_T] : (LbT) 0;

	ElemT key[IPT];
	u32 val[HAS_VAL ? IPT : 1];

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
		const u32* pv = vin + base + wbase;
		if (cnt == (u32) TILE) {
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				key[i] = __ldcs(p + i * 32);
				if (HAS_VAL) val[i] = __ldcs(pv + i * 32);
			}
		} else {
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const bool in_tile = wbase + i * 32u < cnt;
				key[i] = in_tile ? __ldcs(p + i * 32) : ElemT(0);
				if (HAS_VAL) val[i] = in_tile ? __ldcs(pv + i * 32) : 0u;
			}
		}
	};
	auto zero_row = [&]() { reinterpret_cast<uint4*>(wh)[lane] = make_uint4(0u, 0u, 0u, 0u); };
	/* +1 on digit d's packed counter; the returning form gives the counter's old value */
	auto row_count = [&](u32 d) { atomicAdd(&wh[d >> 1], 1u << ((d & 1u) << 4)); };
	auto row_take = [&](u32 d) -> u32 {
		const u32 sh = (d & 1u) << 4;
		return (atomicAdd(&wh[d >> 1], 1u << sh) >> sh) & 0xffffu;
	};
	/* Skewed or clustered digits: lanes of one warp instruction that hit one address serialise in
	 * the atomic unit (tools/ubench_atoms_dup.cu: one cycle per lane, 32 for a constant digit,
	 * with or without a return value once the addend is not the constant 1).  A warp then counts
	 * and places the keys of two CANDIDATE digits with ballots instead (deterministic, stable, no
	 * atomic): the digit of its first key, and either the heaviest digit of the last counted tile
	 * (when that digit holds >= 1/8 of the tile) or the digit of its last key (sorted or clustered
	 * input -- also what every pass after the first sees when keys repeat -- where the 512 keys
	 * of a warp span at most two digits).  A warp takes the path when the tile before was heavy
	 * or when its first 32 keys mostly agree -- and looks at all only when the last counted tile was
	 * LUMPY (some digit with >= 1/64 of the keys, four times the mean; uniform keys never are), so
	 * ordinary tiles keep the plain atomics at the cost of one register compare. */
	auto skew_mode = [&](bool heavy, bool lumpy) -> bool {
		if (!lumpy || (flags & 32)) return false;      /* flag 32: A/B switch, plain atomics whatever the digits are */
		const u32 d0 = v6_digit<ElemT>(key[0], start_bit, dmask);
		return heavy || __popc(__ballot_sync(0xffffffffu, d0 == __shfl_sync(0xffffffffu, d0, 0))) >= 16;
	};
	auto skew_candidates = [&](bool heavy, u32 heavy_digit, u32& h1, u32& h2) {
		h2 = __shfl_sync(0xffffffffu, v6_digit<ElemT>(key[0], start_bit, dmask), 0);
		h1 = heavy ? heavy_digit : __shfl_sync(0xffffffffu, v6_digit<ElemT>(key[IPT - 1], start_bit, dmask), 31);
		if (h1 == h2) h1 = 0x100u;
	};
	/* staging slot -> position: an XOR swizzle of the low five bits, so that the 32 lanes of a
	 * placement hit 32 banks also when their slots are a multiple of 32 apart (sorted input: lane
	 * l places digit d0 + l, and every digit run of the tile is 32 keys long); a linear, 32-aligned
	 * read of the write-out stays conflict free */
	auto sw = [](u32 j) -> u32 { return CLO_NO_SWIZZLE ? j : (j ^ ((j >> 5) & 31u)); };
	/* coalesced write-out of a staged tile; returns true when the staged order is not
	 * sorted on the bits processed so far (= some warp instruction's atomics were not
	 * applied in lane order) */
	/* executed by the threads [first, first + nthr) of the CTA (nthr a multiple of 32); bar_id names
	 * the barrier those threads share (0 = the whole CTA) */
	auto write_out = [&](u32 t, int b, int gslot, bool verify, const int first, const int nthr, const int bar_id) -> bool {
		const u32 cnt = tile_count_of(t);
		const u32 lid = (u32) (tid - first);
		const ElemT* skeys = buf_keys(b);
		const u32* svals = buf_vals(b);
		const unsigned short* sinfo = buf_info(b);
		const LbT* goff = goff_of(gslot);
		bool bad = false;
		if (flags & 16) {
			/* test hook: plant a REAL inversion -- swap two neighbouring staged entries of the same
			 * digit -- before the detector looks, so the predicate itself is exercised */
			if (lid == 0 && (t % 7u) == 3u) {
				ElemT* wk = const_cast<ElemT*>(skeys);
				u32* wv = const_cast<u32*>(svals);
				unsigned short* wi = const_cast<unsigned short*>(sinfo);
				for (u32 j = 1; j < cnt; ++j) {
					const u32 ja = sw(j - 1), jb = sw(j);
					const ElemT a = wk[ja], b2 = wk[jb];
					if (v6_digit<ElemT>(a, start_bit, dmask) != v6_digit<ElemT>(b2, start_bit, dmask)) continue;
					if (!HAS_VAL && (a & low_mask) == (b2 & low_mask)) continue;
					wk[ja] = b2; wk[jb] = a;
					if (HAS_VAL) {
						const u32 v = wv[ja]; wv[ja] = wv[jb]; wv[jb] = v;
						const unsigned short x = wi[ja]; wi[ja] = wi[jb]; wi[jb] = x;
					}
					atomicAdd(err_flag + 2, 1);
					break;
				}
			}
			named_bar_sync(bar_id, nthr);
		}
		if ((flags & 8) && (t % 5u) == 2u) {
			/* test hook: write this tile WRONG and report it, so that only a working repair
			 * path gives a sorted result */
			for (u32 j = lid; j < cnt; j += (u32) nthr) {
				const ElemT k = skeys[sw(j)];
				const LbT o = goff[v6_digit<ElemT>(k, start_bit, dmask)] + (LbT) j;
				out[o] = (ElemT) ~k;
				if (HAS_VAL) vout[o] = ~svals[sw(j)];
			}
			return true;
		}
		auto one = [&](u32 j) {
			const u32 js = sw(j);
			const ElemT k = skeys[js];
			const u32 d = v6_digit<ElemT>(k, start_bit, dmask);
			if (VERIFY) {
				const u32 jp = sw(j > 0 ? j - 1 : 0);
				const ElemT kp = skeys[jp];
				if (HAS_VAL) {
					/* equal digits must keep their order in the tile (payloads tell equal keys apart) */
					if (j > 0 && v6_digit<ElemT>(kp, start_bit, dmask) == d && sinfo[js] <= sinfo[jp]) bad = true;
				} else {
					if ((k & low_mask) < (kp & low_mask)) bad = true;
				}
			}
			const LbT o = goff[d] + (LbT) j;
			out[o] = k;
			if (HAS_VAL) vout[o] = svals[js];
		};
		if (nthr == THREADS) {
			if (cnt == (u32) TILE) {
#pragma unroll
				for (int i = 0; i < IPT; ++i) one(lid + i * THREADS);
			} else {
#pragma unroll
				for (int i = 0; i < IPT; ++i) {
					const u32 j = lid + i * THREADS;
					if (j < cnt) one(j);
				}
			}
		} else {
			/* the write-out that runs beside the digit phase: THREADS - DT threads */
			constexpr int WT = THREADS - DT, FULL = TILE / WT, REM = TILE - FULL * WT;
			if (cnt == (u32) TILE) {
#pragma unroll
				for (int i = 0; i < FULL; ++i) one(lid + i * (u32) WT);
				if (REM > 0 && lid < (u32) REM) one(lid + FULL * (u32) WT);
			} else {
#pragma unroll 1
				for (u32 j = lid; j < cnt; j += (u32) WT) one(j);
			}
		}
		return verify && bad;
	};
	/* prefix threads: PREF[t] -> global offset table of the tile staged in buffer b */
	auto prefix_to_goff = [&](u32 t, int gslot, int dslot, const LbT (&w0)[DPT]) {
#pragma unroll
		for (int j = 0; j < DPT; ++j) {
			const int d = tid - DT + j * PREF_T;
			LbT* p = pref + (size_t) t * RADIX + d;
			LbT w = w0[j];
			unsigned spins = 0;
			while (!(w & PPWord<LbT>::VALID)) {
				if (++spins > SPIN_LIMIT) { atomicExch(err_flag, 1); break; }
				w = ld_relaxed(p);
			}
			if (prof_on && d == 0) atomicAdd(prof + 8, (u64) spins);
			st_relaxed(p, (LbT) 0);                                  /* consumed: reset */
			goff_of(gslot)[d] = my_base[j] + (w & PPWord<LbT>::VAL) - (LbT) s_ds[dslot * RADIX + d];
		}
	};
	auto repair = [&](u32 t, int b, int gslot, int dslot) {
		v6_repair<ElemT, LbT, THREADS, IPT, HAS_VAL>(in, out, vin, vout, n, t, reinterpret_cast<u32*>(buf_keys(b)),
			goff_of(gslot), s_ds + dslot * RADIX, start_bit, dmask, err_flag);
	};

	/* test hook: pretend a prefix wait timed out, so the status reporting of the host side can be
	 * exercised (tests/test_gpu_sort.py::test_satradix_timeout_is_reported) */
	if ((flags & 64) && blockIdx.x == V6_NUM_PROP && tid == 0) atomicExch(err_flag, 1);
	/* ---- prologue: two tickets (the next tile is always known one iteration ahead) */
	if (tid == 0) { s_misc[8] = atomicAdd(ticket, 1u); s_misc[9] = atomicAdd(ticket, 1u); s_misc[10] = 0; s_misc[11] = 0; }
	zero_row();
	__syncthreads();
	u32 cur = s_misc[8];
	u32 nxt = s_misc[9];
	if (cur >= num_tiles) return;
	__syncthreads();
	load_tile(cur);
	{
	/* ---- schedule, TWO iterations of slack between a tile's AGG and the use of its PREF, two
	 * CTA barriers per tile:
	 *   P1 count(k) | B1 | warps 0-3: digits(k) -> AGG, bases  ||  the other warps: PREF(k-2) ->
	 *   offsets, then write-out(k-2) | B2 | place(k) into the buffer just freed + load(k+1)
	 * The latency-bound digit phase (128 threads) runs UNDER the write-out of the older tile
	 * instead of in front of it.  The prefix latency (propagator rounds, L2 round trips, slow
	 * neighbours) is far from the critical path and the propagators need no SM of their own.
	 * Tickets are drawn two tiles ahead, so the next tile's keys are pulled into L2 a whole
	 * iteration before they are loaded. */
	u32 t1 = NONE, t2 = NONE;         /* tiles of the previous two iterations (staged, not yet written) */
	int b = 0;                        /* buffer of the current tile (= buffer of t2) */
	int s0 = 0, s1 = 2, s2 = 1;       /* s_ds slots of cur, t1, t2 (k, k-1, k-2 mod 3) */
	const LbT zero_w[DPT] = {};
	constexpr int WO_T = THREADS - DT;        /* threads of the overlapped write-out */
	/* a tile is HEAVY when one digit holds >= TILE >> hv_shift keys (1/8; flags bits 8-9: 1/16, 1/4, 1/32 for A/B) */
	const int hv_shift = ((flags >> 8) & 3) == 0 ? 3 : (((flags >> 8) & 3) == 1 ? 4 : (((flags >> 8) & 3) == 2 ? 2 : 5));
	bool lumpy = false;                       /* the last counted tile has a digit with >= 1/64 of its keys */
	bool heavy = false;                       /* ... with >= 1/8 of its keys ... */
	u32 heavy_digit = 0;                      /* ... this one */
	for (;;) {
		LbT wp[DPT] = {};
		if (t2 != NONE && is_pref_thread) {
#pragma unroll
			for (int j = 0; j < DPT; ++j) wp[j] = ld_relaxed(pref + (size_t) t2 * RADIX + (tid - DT + j * PREF_T));
		}
		const u32 cnt = tile_count_of(cur);
		/* P1 count */
		if (cnt == (u32) TILE && skew_mode(heavy, lumpy)) {
			u32 h1, h2, c1 = 0, c2 = 0;
			skew_candidates(heavy, heavy_digit, h1, h2);
#pragma unroll
			for (int i = 0; i < IPT; ++i) {
				const u32 d = v6_digit<ElemT>(key[i], start_bit, dmask);
				c1 += __popc(__ballot_sync(0xffffffffu, d == h1));
				c2 += __popc(__ballot_sync(0xffffffffu, d == h2));
				if (d != h1 && d != h2) row_count(d);
			}
			if (lane == 0) {
				if (c1) atomicAdd(&wh[h1 >> 1], c1 << ((h1 & 1u) << 4));
				if (c2) atomicAdd(&wh[h2 >> 1], c2 << ((h2 & 1u) << 4));
			}
		} else if (cnt == (u32) TILE) {
#pragma unroll
			for (int i = 0; i < IPT; ++i) row_count(v6_digit<ElemT>(key[i], start_bit, dmask));
		} else {
#pragma unroll
			for (int i = 0; i < IPT; ++i)
				if (wbase + i * 32u < cnt) row_count(v6_digit<ElemT>(key[i], start_bit, dmask));
		}
		mark(0);
		__syncthreads();                                         /* B1 */
		mark(1);
		bool bad = false;
		if (tid < DT) {
			/* P2: my two digits, 2 * tid (low halves) and 2 * tid + 1 (high halves) */
			u32 c[WARPS];
#pragma unroll
			for (int w = 0; w < WARPS; ++w) c[w] = whist[w * ROWW + tid];
			u32 sum = 0;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) sum += c[w];
			const u32 c0 = sum & 0xffffu, c1 = sum >> 16;
			{
				/* the tile's heaviest digit, (count << 8) | digit: selects the skew paths of the
				 * placement below and of the next tile's count */
				u32 top = c1 > c0 ? ((c1 << 8) | (u32) (2 * tid + 1)) : ((c0 << 8) | (u32) (2 * tid));
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) top = max(top, __shfl_xor_sync(0xffffffffu, top, o));
				if (lane == 0) s_misc[12 + warp] = top;
			}
			st_relaxed(agg + (size_t) cur * RADIX + 2 * tid, (LbT) (PPWord<LbT>::VALID | (LbT) c0));
			st_relaxed(agg + (size_t) cur * RADIX + 2 * tid + 1, (LbT) (PPWord<LbT>::VALID | (LbT) c1));
			const u32 pair = c0 + c1;
			const u32 incl = warp_inclusive_scan<u32>(pair, lane);
			if (lane == 31) s_misc[warp] = incl;
			named_bar_sync(1, DT);
			u32 off = 0;
#pragma unroll
			for (int w = 0; w < DT / 32; ++w) if (w < warp) off += s_misc[w];
			if (tid == 0) s_misc[16] = max(max(s_misc[12], s_misc[13]), max(s_misc[14], s_misc[15]));
			const u32 ds0 = off + incl - pair, ds1 = ds0 + c0;
			*reinterpret_cast<uint2*>(s_ds + s0 * RADIX + 2 * tid) = make_uint2(ds0, ds1);
			u32 run = ds0 | (ds1 << 16);
#pragma unroll
			for (int w = 0; w < WARPS; ++w) { whist[w * ROWW + tid] = run; run += c[w]; }
			mark(2);
		} else {
			if (is_pref_thread) {
				u32 nt = 0;
				if (tid == DT) nt = atomicAdd(ticket, 1u);           /* the tile after next; in flight during the prefix wait */
				if (t2 != NONE) prefix_to_goff(t2, 0, s2, wp);
				if (tid == DT) s_misc[8] = nt;
			}
			/* P5 write-out of the tile staged two iterations ago (same buffer as the current tile) */
			if (t2 != NONE) {
				named_bar_sync(2, WO_T);
				bad = write_out(t2, b, 0, true, DT, WO_T, 2);
			}
		}
		if (__syncthreads_or(bad ? 1 : 0)) repair(t2, b, 0, s2);   /* B2 */
		mark(3);
		const u32 nxt2 = s_misc[8];
		const bool more = nxt < num_tiles;
		heavy = (s_misc[16] >> 8) >= ((u32) TILE >> hv_shift);
		lumpy = (s_misc[16] >> 8) >= (u32) (TILE / 64);
		heavy_digit = s_misc[16] & 0xffu;
		if (nxt2 < num_tiles) {
			/* the tile after next -> L2 (one 128-byte line per thread) */
			const size_t lo = (size_t) nxt2 * TILE * sizeof(ElemT) + (size_t) tid * 128;
			if (lo < n * sizeof(ElemT) && (size_t) tid * 128 < (size_t) TILE * sizeof(ElemT))
				asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(in) + lo));
			if (HAS_VAL) {
				const size_t lv = (size_t) nxt2 * TILE * 4 + (size_t) tid * 128;
				if (lv < n * 4 && (size_t) tid * 128 < (size_t) TILE * 4)
					asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(vin) + lv));
			}
		}
		/* P3 place + P4 next tile: a key register is refilled as soon as its key is staged */
		{
			ElemT* skeys = buf_keys(b);
			u32* svals = buf_vals(b);
			unsigned short* sinfo = buf_info(b);
			const bool next_full = more && tile_count_of(nxt) == (u32) TILE;
			if (cnt == (u32) TILE && skew_mode(heavy, lumpy)) {
				/* skew path: keys of the candidate digits take their slots from a ballot and a
				 * warp-private running base; the rest as usual.  (The next tile is loaded after
				 * the loop: refilling a register right behind a ballot made ptxas wait for each load.) */
				u32 h1, h2;
				skew_candidates(heavy, heavy_digit, h1, h2);
				u32 b1 = h1 < 0x100u ? ((wh[h1 >> 1] >> ((h1 & 1u) << 4)) & 0xffffu) : 0u;
				u32 b2 = (wh[h2 >> 1] >> ((h2 & 1u) << 4)) & 0xffffu;
				const u32 lt = lanemask_lt();
#pragma unroll
				for (int i = 0; i < IPT; ++i) {
					const u32 d = v6_digit<ElemT>(key[i], start_bit, dmask);
					const u32 m1 = __ballot_sync(0xffffffffu, d == h1), m2 = __ballot_sync(0xffffffffu, d == h2);
					u32 p;
					if (d == h1) p = b1 + __popc(m1 & lt);
					else if (d == h2) p = b2 + __popc(m2 & lt);
					else p = row_take(d);
					b1 += __popc(m1); b2 += __popc(m2);
					p = sw(p);
					skeys[p] = key[i];
					if (HAS_VAL) {
						svals[p] = val[i];
						sinfo[p] = (unsigned short) (wbase + i * 32u);
					}
				}
				__syncwarp();
				zero_row();
				if (more) load_tile(nxt);
			} else if (cnt == (u32) TILE && next_full) {
				const ElemT* np = in + (size_t) nxt * TILE + wbase;
				const u32* npv = vin + (size_t) nxt * TILE + wbase;
#pragma unroll
				for (int i = 0; i < IPT; ++i) {
					const u32 p = sw(row_take(v6_digit<ElemT>(key[i], start_bit, dmask)));
					skeys[p] = key[i];
					key[i] = __ldcs(np + i * 32);
					if (HAS_VAL) {
						svals[p] = val[i];
						sinfo[p] = (unsigned short) (wbase + i * 32u);
						val[i] = __ldcs(npv + i * 32);
					}
				}
				__syncwarp();
				zero_row();
			} else {
#pragma unroll
				for (int i = 0; i < IPT; ++i)
					if (wbase + i * 32u < cnt) {
						const u32 p = sw(row_take(v6_digit<ElemT>(key[i], start_bit, dmask)));
						skeys[p] = key[i];
						if (HAS_VAL) {
							svals[p] = val[i];
							sinfo[p] = (unsigned short) (wbase + i * 32u);
						}
					}
				__syncwarp();
				zero_row();
				if (more) load_tile(nxt);
			}
		}
		mark(6);
		t2 = t1; t1 = cur;
		b ^= 1;
		{ const int t = s2; s2 = s1; s1 = s0; s0 = t; }
		if (!more) break;
		cur = nxt;
		nxt = nxt2;
	}
	/* ---- epilogue: t2 is staged in buffer b, t1 in buffer b^1 */
	for (int r = 0; r < 2; ++r) {
		const u32 t = r == 0 ? t2 : t1;
		const int bb = r == 0 ? b : (b ^ 1);
		const int ss = r == 0 ? s2 : s1;
		__syncthreads();
		if (t == NONE) continue;
		if (is_pref_thread) prefix_to_goff(t, 0, ss, zero_w);
		__syncthreads();
		const bool bad = write_out(t, bb, 0, true, 0, THREADS, 0);
		if (__syncthreads_or(bad ? 1 : 0)) repair(t, bb, 0, ss);
	}
	}
}

/* Are same-address shared-memory atomics of one warp instruction served in lane order?  The
 * instruction form of the placement above (packed 16-bit counters, add with return) on
 * addresses that collide from "sometimes" to "always"; out[0] counts violations against the
 * __match_any_sync ranks, out[1] the samples. */
__global__ void clo_radix_atomic_order_selftest_kernel(u32* __restrict__ out, u32 seed, int iters) {
	__shared__ u32 row[16][RADIX / 2];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	u32 s = seed + threadIdx.x * 7919u + blockIdx.x * 104729u;
	u32 bad = 0, tot = 0;
	for (int it = 0; it < iters; ++it) {
		reinterpret_cast<uint4*>(row[warp])[lane] = make_uint4(0u, 0u, 0u, 0u);
		__syncwarp();
		s = s * 1664525u + 1013904223u;
		const u32 spread = (it & 3) == 0 ? 255u : ((it & 3) == 1 ? 15u : ((it & 3) == 2 ? 3u : 0u));
		const u32 d = (s >> 24) & spread;
		const u32 sh = (d & 1u) << 4;
		const u32 old = (atomicAdd(&row[warp][d >> 1], 1u << sh) >> sh) & 0xffffu;
		const u32 peers = __match_any_sync(0xffffffffu, d);
		bad += (old != (u32) __popc(peers & ((1u << lane) - 1u)));
		tot += 1;
		__syncwarp();
	}
	atomicAdd(out, bad);
	atomicAdd(out + 1, tot);
}

template <typename ElemT, int THREADS, int IPT, typename LbT, bool HAS_VAL = false>
constexpr size_t onesweep_v6_smem() {
	constexpr size_t worker = (size_t) (THREADS / 32) * (RADIX / 2) * 4 + 4 * RADIX * 4 + 2 * RADIX * 8 + 32 * 4 +
		2 * ((size_t) THREADS * IPT * sizeof(ElemT) + (HAS_VAL ? (size_t) THREADS * IPT * 6 : 0));
	constexpr size_t prop = pp_propagate2_smem<LbT, THREADS, v6_prop_groups(THREADS, (int) sizeof(LbT))>();
	return worker > prop ? worker : prop;
}

#endif
