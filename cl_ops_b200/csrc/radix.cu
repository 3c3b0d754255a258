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

struct PassCfg {
	int passes;
	u32 start_bit[MAX_PASSES];
	u32 dmask[MAX_PASSES];
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
		/* four vectors in flight per thread */
		for (; i + 3 * stride < nvec; i += 4 * stride) {
			ElemT e[4][EPV];
#pragma unroll
			for (int u = 0; u < 4; ++u) load_vec_cs<ElemT, EPV>(in + (i + u * stride) * EPV, e[u]);
#pragma unroll
			for (int u = 0; u < 4; ++u)
#pragma unroll
				for (int c = 0; c < EPV; ++c) count(e[u][c]);
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

/* (the one-tile-per-CTA look-back kernel of round 1 and its splitter-as-digit partition variant
 * were removed in round 2: every configuration runs a persistent kernel, radix_pp.cuh or
 * radix_v6.cuh) */

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

template <typename ElemT, bool HAS_VAL, bool IDENTITY, int THREADS, int IPT>
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
	/* tile-prefix words keep 31 value bits (every prefix and offset is a count of keys, i.e. < n) */
	const bool wide = st->force_wide || n >= (1ull << 31);
	WorkLayout L;
	if ((e = prepare_work(st, 0, cfg.passes, wide ? 8 : 4, L, stream)) != cudaSuccess) return e;
	{
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
	/* chain mode (every pass runs the v6 kernel): a pass whose digit is the same for every key --
	 * small keys, sorted prefixes, all-equal input -- moves nothing; the passes hand the location of
	 * the keys on through 16-byte chain words and one conditional copy at the end puts the result
	 * into dst.  Flag 128 keeps the round-1 behaviour (identity pass = copy) for A/B. */
	bool chain_mode = false;
	if constexpr (IDENTITY && sizeof(ElemT) >= 4 && THREADS > RADIX) chain_mode = st->kernel_v6 && !(g_pp_flags & 128);
	char* chain = (char*) L.tickets + 64;                     /* [MAX_PASSES + 1] chain words behind the tickets */
	const ElemT* cur = src; const u32* vcur = vsrc;
	if (!chain_mode && (const void*) src == (const void*) dst && (cfg.passes & 1)) {
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
			if (st->kernel_v6) {
				char* agg = (char*) st->pp.ptr;
				char* pref = agg + tiles * RADIX * (wide ? 8 : 4);
				if (chain_mode) {
					ElemT* alt = to_dst ? aux : dst;
					u32* valt = to_dst ? vaux : vdst;
					e = clo_radix_v6_pass((int) sizeof(ElemT), wide ? 1 : 0, THREADS * IPT, src, nxt, HAS_VAL ? vsrc : nullptr, HAS_VAL ? vnxt : nullptr, n, agg, pref, ticket,
						L.bins + p * RADIX, cfg.start_bit[p], cfg.dmask[p], L.err, sm_count, g_radix_profile, g_pp_flags, stream,
						p == 0 ? nullptr : chain + 16 * p, chain + 16 * (p + 1), alt, HAS_VAL ? valt : nullptr);
				} else
				e = clo_radix_v6_pass((int) sizeof(ElemT), wide ? 1 : 0, THREADS * IPT, cur, nxt, HAS_VAL ? vcur : nullptr, HAS_VAL ? vnxt : nullptr, n, agg, pref, ticket,
					L.bins + p * RADIX, cfg.start_bit[p], cfg.dmask[p], L.err, sm_count, g_radix_profile, g_pp_flags, stream);
				done_v6 = true;
			}
		}
		if (done_v6) {
		} else {
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
		}
		if (e != cudaSuccess) return e;
		st->mark(stream);
		cur = nxt; vcur = vnxt;
	}
	if (chain_mode && cfg.passes > 0)
		return clo_radix_v6_chain_fixup(chain + 16 * cfg.passes, dst, HAS_VAL ? (void*) vdst : nullptr, n * sizeof(ElemT), n * sizeof(u32), sm_count, stream);
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
	if (n >= (1ull << 32)) { if (err_msg) *err_msg = "partition: at most 2^32 - 1 elements per GPU"; return cudaErrorInvalidValue; }
	if (elem_size == 4 || elem_size == 8) {
		/* the chunk-per-warp partition (partition.cu) */
		int dev = 0, sms = 0;
		cudaGetDevice(&dev);
		if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
		return clo_partition_v2(st->work, elem_size, keys_in, payload_in, keys_out, payload_out, n, gidx0,
			splitter_keys, splitter_idx, nparts, counts_out, sms, stream, err_msg);
	}
	if (err_msg) *err_msg = "partition: keys must be 4 or 8 bytes";
	return cudaErrorInvalidValue;
}
