/*
 * rng.cu -- clo_rng: device seeds buffer + the six generators as vectorised
 * per-thread bulk generators (replaces /root/reference/src/cl_ops/rng/clo_rng.c
 * and the OpenCL C device functions clo_rng_{lcg,xorshift64,xorshift128,mwc64x,
 * parkmiller,tauslcg}.cl, clo_rng_init.cl, clo_rng_api.cl).
 *
 * The reference keeps generator state in global memory and round-trips it on
 * every number (e.g. clo_rng_lcg.cl:49-55).  Here a thread owns 4 consecutive
 * streams, keeps their states in registers for all `runs`, and writes one
 * 16-byte store per run; states go back to memory once, so a later call
 * continues every stream bit-exactly.
 *
 * HBM traffic: 4 bytes per generated word (+ 2 * seed_size per stream per call).
 */
#include "clo_internal.h"
#include <string>
#include <cctype>
#include "device_utils.cuh"

#include <cstring>

using namespace clo;

namespace {

enum { R_LCG = 0, R_XS64 = 1, R_XS128 = 2, R_MWC64X = 3, R_PARKMILLER = 4, R_TAUSLCG = 5, R_COUNT = 6 };
enum { H_NONE = 0, H_KNUTH = 1, H_XS1 = 2 };

/* clo_rng_init.cl:29-35, 64-bit ulong arithmetic */
__device__ __forceinline__ u64 seed_hash(int hash, u64 x) {
	if (hash == H_KNUTH) {
		x = (x * 2654435761ull) % 0x100000000ull;
	} else if (hash == H_XS1) {
		x = ((x >> 16) ^ x) * 0x45d9f3bull;
		x = ((x >> 16) ^ x) * 0x45d9f3bull;
		x = ((x >> 16) ^ x);
	}
	return x;
}

template <int R> struct Gen;

/* clo_rng_lcg.cl:29-59 */
template <> struct Gen<R_LCG> {
	typedef u64 State;
	__device__ static State from_seed(u64 s) { return s; }
	__device__ static u32 next(State& s) {
		s = (s * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1);
		return (u32) (s >> 16);
	}
};

/* clo_rng_xorshift64.cl:26-63 */
template <> struct Gen<R_XS64> {
	typedef u64 State;
	__device__ static State from_seed(u64 s) { return s; }
	__device__ static u32 next(State& s) {
		s ^= (s << 21); s ^= (s >> 35); s ^= (s << 4);
		return (u32) s;
	}
};

/* clo_rng_xorshift128.cl:27-59 (the fourth component really is seed >> 46) */
template <> struct Gen<R_XS128> {
	typedef uint4 State;
	__device__ static State from_seed(u64 s) {
		return make_uint4((u32) s, (u32) (s >> 16), (u32) (s >> 32), (u32) (s >> 46));
	}
	__device__ static u32 next(State& s) {
		u32 t = s.x ^ (s.x << 11);
		s.x = s.y; s.y = s.z; s.z = s.w;
		s.w = s.w ^ (s.w >> 19) ^ (t ^ (t >> 8));
		return s.w;
	}
};

/* clo_rng_mwc64x.cl:28-63 */
template <> struct Gen<R_MWC64X> {
	typedef uint2 State;
	__device__ static State from_seed(u64 s) { return make_uint2((u32) s, (u32) (s >> 32)); }
	__device__ static u32 next(State& s) {
		const u32 A = 4294883355u;
		u32 x = s.x, c = s.y;
		u32 res = x ^ c;
		u32 hi = __umulhi(x, A);
		x = x * A + c;
		c = hi + (x < c);
		s.x = x; s.y = c;
		return res;
	}
};

/* clo_rng_parkmiller.cl:28-59: (long) state * 16807 % INT_MAX with C's
 * truncating remainder, so a negative state stays negative. */
template <> struct Gen<R_PARKMILLER> {
	typedef int State;
	__device__ static State from_seed(u64 s) { return (int) (u32) s; }
	__device__ static u32 next(State& s) {
		/* |s| * 16807 < 2^46; 2^31 = 1 (mod 2^31-1), so two folds of the high bits replace the
		 * 64-bit division; the truncating remainder takes the sign of the dividend */
		const u32 a = s < 0 ? 0u - (u32) s : (u32) s;
		const u32 lo = a * 16807u, hi = __umulhi(a, 16807u);
		u32 r = (lo & 0x7fffffffu) + ((hi << 1) | (lo >> 31));
		r = (r & 0x7fffffffu) + (r >> 31);
		if (r >= 0x7fffffffu) r -= 0x7fffffffu;
		s = s < 0 ? -(int) r : (int) r;
		return ((u32) s) << 1;
	}
};

/* clo_rng_tauslcg.cl:32-100 */
__device__ __forceinline__ u32 taus_step(u32 z, int s1, int s2, int s3, u32 m) {
	u32 b = (((z << s1) ^ z) >> s2);
	return (((z & m) << s3) ^ b);
}

template <> struct Gen<R_TAUSLCG> {
	typedef uint4 State;
	__device__ static State from_seed(u64 s) {
		return make_uint4((u32) s, (u32) (s >> 32), (u32) s, (u32) (s >> 32));
	}
	__device__ static u32 next(State& s) {
		u32 x = s.x;
		s.x = taus_step(s.y, 13, 19, 12, 4294967294u);
		s.y = taus_step(s.z, 2, 25, 4, 4294967288u);
		s.z = taus_step(s.w, 3, 11, 17, 4294967294u);
		s.w = 1664525u * x + 1013904223u;
		return s.x;
	}
};

/* clo_rng_init.cl:47-60 */
template <int R>
__global__ void clo_rng_init_kernel(typename Gen<R>::State* __restrict__ states, size_t count,
		u64 gid0, u64 main_seed, int hash) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	u64 seed = (gid0 + i) + main_seed;
	seed = seed_hash(hash, seed);
	states[i] = Gen<R>::from_seed(seed);
}

/* states from already hashed 64-bit seeds (custom CLO_RNG_HASH strings, built by jit.cu) */
template <int R>
__global__ void clo_rng_init_from_seeds_kernel(typename Gen<R>::State* __restrict__ states, size_t count,
		const u64* __restrict__ seeds) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count) states[i] = Gen<R>::from_seed(seeds[i]);
}

/* Bulk generation, layout of clo_rng_bench.cl:23-37 + clo_rng_bench.c:302-324:
 * out[r * G + g] = f(next_r(state_g)), f = >> shift  or  % maxint. */
template <int R>
__global__ void __launch_bounds__(256)
clo_rng_generate_kernel(typename Gen<R>::State* __restrict__ states, size_t G, size_t runs,
		u32* __restrict__ out, u32 shift, u32 maxint, int vec_ok) {
	typedef typename Gen<R>::State State;
	const size_t g0 = ((size_t) blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (g0 >= G) return;
	State s[4];
	const int live = (G - g0 >= 4) ? 4 : (int) (G - g0);
#pragma unroll
	for (int k = 0; k < 4; ++k) if (k < live) s[k] = states[g0 + k];
	if (live == 4 && vec_ok) {
		uint4* o = reinterpret_cast<uint4*>(out + g0);
		const size_t stride = G / 4;
		if (maxint == 0) {
			for (size_t r = 0; r < runs; ++r) {
				uint4 w;
				w.x = Gen<R>::next(s[0]) >> shift;
				w.y = Gen<R>::next(s[1]) >> shift;
				w.z = Gen<R>::next(s[2]) >> shift;
				w.w = Gen<R>::next(s[3]) >> shift;
				__stcs(o, w);
				o += stride;
			}
		} else {
			for (size_t r = 0; r < runs; ++r) {
				uint4 w;
				w.x = Gen<R>::next(s[0]) % maxint;
				w.y = Gen<R>::next(s[1]) % maxint;
				w.z = Gen<R>::next(s[2]) % maxint;
				w.w = Gen<R>::next(s[3]) % maxint;
				__stcs(o, w);
				o += stride;
			}
		}
	} else {
		for (size_t r = 0; r < runs; ++r) {
#pragma unroll
			for (int k = 0; k < 4; ++k) {
				if (k < live) {
					u32 v = Gen<R>::next(s[k]);
					out[r * G + g0 + k] = maxint ? (v % maxint) : (v >> shift);
				}
			}
		}
	}
#pragma unroll
	for (int k = 0; k < 4; ++k) if (k < live) states[g0 + k] = s[k];
}

template <int R>
cudaError_t launch_init(void* states, size_t count, u64 gid0, u64 main_seed, int hash, cudaStream_t stream) {
	if (!count) return cudaSuccess;
	const unsigned blocks = (unsigned) ((count + 255) / 256);
	clo_rng_init_kernel<R><<<blocks, 256, 0, stream>>>((typename Gen<R>::State*) states, count, gid0, main_seed, hash);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

template <int R>
cudaError_t launch_generate(void* states, size_t G, size_t runs, void* out, u32 bits, u32 maxint, cudaStream_t stream) {
	if (!G || !runs) return cudaSuccess;
	const size_t threads = (G + 3) / 4;
	const unsigned blocks = (unsigned) ((threads + 255) / 256);
	const int vec_ok = (G % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
	clo_rng_generate_kernel<R><<<blocks, 256, 0, stream>>>((typename Gen<R>::State*) states, G, runs,
		(u32*) out, 32u - bits, maxint, vec_ok);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

typedef cudaError_t (*InitFn)(void*, size_t, u64, u64, int, cudaStream_t);
template <int R>
cudaError_t launch_init_from_seeds(void* states, size_t count, const u64* seeds, cudaStream_t stream) {
	if (!count) return cudaSuccess;
	clo_rng_init_from_seeds_kernel<R><<<(unsigned) ((count + 255) / 256), 256, 0, stream>>>((typename Gen<R>::State*) states, count, seeds);
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}
typedef cudaError_t (*InitSeedsFn)(void*, size_t, const u64*, cudaStream_t);

typedef cudaError_t (*GenFn)(void*, size_t, size_t, void*, u32, u32, cudaStream_t);

const InitFn kInit[R_COUNT] = { launch_init<0>, launch_init<1>, launch_init<2>, launch_init<3>, launch_init<4>, launch_init<5> };
const InitSeedsFn kInitSeeds[R_COUNT] = { launch_init_from_seeds<0>, launch_init_from_seeds<1>, launch_init_from_seeds<2>,
	launch_init_from_seeds<3>, launch_init_from_seeds<4>, launch_init_from_seeds<5> };
const GenFn kGen[R_COUNT] = { launch_generate<0>, launch_generate<1>, launch_generate<2>, launch_generate<3>, launch_generate<4>, launch_generate<5> };

/* ------------------------------------------------------------- sources */
/* CUDA device source of each generator, the analogue of the OpenCL source the
 * reference hands to its clients (clo_rng.c:371-372): clo_statetype,
 * clo_rng_next(states, index) and the next_int{,2,4,8} API of clo_rng_api.cl:33-105
 * (work-item k of a vector call uses the streams gid + k * global_size, clo_rng_workitem.cl:24-32). */

#define CLO_SRC_API \
	"__device__ inline unsigned clo_rng_next_int(clo_statetype* states, unsigned n) {\n" \
	"\tunsigned index = blockIdx.x * blockDim.x + threadIdx.x;\n" \
	"\treturn clo_rng_next(states, index) % n;\n}\n" \
	"__device__ inline uint2 clo_rng_next_int2(clo_statetype* states, unsigned n) {\n" \
	"\tunsigned g = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;\n" \
	"\treturn make_uint2(clo_rng_next(states, g) % n, clo_rng_next(states, gs + g) % n);\n}\n" \
	"__device__ inline uint4 clo_rng_next_int4(clo_statetype* states, unsigned n) {\n" \
	"\tunsigned g = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;\n" \
	"\treturn make_uint4(clo_rng_next(states, g) % n, clo_rng_next(states, gs + g) % n,\n" \
	"\t\tclo_rng_next(states, 2 * gs + g) % n, clo_rng_next(states, 3 * gs + g) % n);\n}\n" \
	"struct uint8 { unsigned s0, s1, s2, s3, s4, s5, s6, s7; };\n" \
	"__device__ inline uint8 clo_rng_next_int8(clo_statetype* states, unsigned n) {\n" \
	"\tunsigned g = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;\n" \
	"\tuint8 r;\n" \
	"\tr.s0 = clo_rng_next(states, g) % n; r.s1 = clo_rng_next(states, gs + g) % n;\n" \
	"\tr.s2 = clo_rng_next(states, 2 * gs + g) % n; r.s3 = clo_rng_next(states, 3 * gs + g) % n;\n" \
	"\tr.s4 = clo_rng_next(states, 4 * gs + g) % n; r.s5 = clo_rng_next(states, 5 * gs + g) % n;\n" \
	"\tr.s6 = clo_rng_next(states, 6 * gs + g) % n; r.s7 = clo_rng_next(states, 7 * gs + g) % n;\n" \
	"\treturn r;\n}\n"

const char kSrcLcg[] =
	"typedef unsigned long long clo_statetype;\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tclo_statetype s = states[index];\n"
	"\ts = (s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);\n"
	"\tstates[index] = s;\n\treturn (unsigned) (s >> 16);\n}\n" CLO_SRC_API;
const char kSrcXs64[] =
	"typedef unsigned long long clo_statetype;\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tclo_statetype s = states[index];\n"
	"\ts ^= (s << 21); s ^= (s >> 35); s ^= (s << 4);\n"
	"\tstates[index] = s;\n\treturn (unsigned) s;\n}\n" CLO_SRC_API;
const char kSrcXs128[] =
	"typedef uint4 clo_statetype;\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tclo_statetype s = states[index];\n"
	"\tunsigned t = s.x ^ (s.x << 11);\n"
	"\ts.x = s.y; s.y = s.z; s.z = s.w;\n"
	"\ts.w = s.w ^ (s.w >> 19) ^ (t ^ (t >> 8));\n"
	"\tstates[index] = s;\n\treturn s.w;\n}\n" CLO_SRC_API;
const char kSrcMwc64x[] =
	"typedef uint2 clo_statetype;\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tconst unsigned A = 4294883355U;\n"
	"\tunsigned x = states[index].x, c = states[index].y;\n"
	"\tunsigned res = x ^ c;\n"
	"\tunsigned hi = __umulhi(x, A);\n"
	"\tx = x * A + c;\n\tc = hi + (x < c);\n"
	"\tstates[index] = make_uint2(x, c);\n\treturn res;\n}\n" CLO_SRC_API;
const char kSrcParkMiller[] =
	"typedef int clo_statetype;\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tint s = states[index];\n"
	"\ts = (int) ((((long long) s) * 16807) % 2147483647);\n"
	"\tstates[index] = s;\n\treturn ((unsigned) s) << 1;\n}\n" CLO_SRC_API;
const char kSrcTausLcg[] =
	"typedef uint4 clo_statetype;\n"
	"__device__ inline unsigned clo_taus_step(unsigned z, int s1, int s2, int s3, unsigned m) {\n"
	"\tunsigned b = (((z << s1) ^ z) >> s2);\n\treturn (((z & m) << s3) ^ b);\n}\n"
	"__device__ inline unsigned clo_rng_next(clo_statetype* states, unsigned index) {\n"
	"\tclo_statetype s = states[index];\n\tunsigned x = s.x;\n"
	"\ts.x = clo_taus_step(s.y, 13, 19, 12, 4294967294U);\n"
	"\ts.y = clo_taus_step(s.z, 2, 25, 4, 4294967288U);\n"
	"\ts.z = clo_taus_step(s.w, 3, 11, 17, 4294967294U);\n"
	"\ts.w = 1664525U * x + 1013904223U;\n"
	"\tstates[index] = s;\n\treturn s.x;\n}\n" CLO_SRC_API;

} // namespace

/* clo_rng.c:60-68 */
extern "C" const struct clo_rng_info clo_rng_infos[] = {
	{"lcg", kSrcLcg, 8},
	{"xorshift64", kSrcXs64, 8},
	{"xorshift128", kSrcXs128, 16},
	{"mwc64x", kSrcMwc64x, 8},
	{"parkmiller", kSrcParkMiller, 4},
	{"tauslcg", kSrcTausLcg, 16},
	{NULL, NULL, 0}
};

struct clo_rng {
	int id;
	const char* src;
	CCLBuffer* seeds_device;
	size_t size_in_device;
	size_t seeds_count;
	CCLContext* ctx;
};

/* MT19937 as GLib's GRand draws it (third party; restated from GLib's grand.c,
 * "2.2" seeding): seeds for CLO_RNG_SEED_HOST_MT, clo_rng.c:185-203. */
namespace {
struct HostMT {
	uint32_t mt[624]; int mti;
	explicit HostMT(uint32_t seed) {
		mt[0] = seed;
		for (mti = 1; mti < 624; mti++)
			mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t) mti;
	}
	uint32_t next() {
		static const uint32_t mag01[2] = { 0u, 0x9908b0dfu };
		uint32_t y;
		if (mti >= 624) {
			int kk;
			for (kk = 0; kk < 624 - 397; kk++) {
				y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
				mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1];
			}
			for (; kk < 623; kk++) {
				y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
				mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1];
			}
			y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
			mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1];
			mti = 0;
		}
		y = mt[mti++];
		y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
		return y;
	}
};

const int H_CUSTOM = 100;     /* any other string: built at run time (jit.cu), as the reference builds it */

int parse_hash(const char* hash, GError** err) {
	(void) err;
	if (!hash || !*hash || strcmp(hash, "x") == 0) return H_NONE;
	std::string sq;
	for (const char* p = hash; *p; ++p) if (!isspace((unsigned char) *p)) sq.push_back(*p);
	if (sq == "KNUTH(x)" || sq == "KNUTH(x);") return H_KNUTH;
	if (sq == "XS1(x)" || sq == "XS1(x);") return H_XS1;
	/* an expression that does not assign to x is a no-op statement in the
	 * reference's `CLO_RNG_HASH(seed);` (clo_rng_init.cl:55), e.g. test_rng.c:42 */
	bool assigns = strstr(hash, "KNUTH") || strstr(hash, "XS1") || strstr(hash, "++") || strstr(hash, "--");
	for (const char* p = hash; *p; ++p)
		if (*p == '=' && p[1] != '=' && (p == hash || (p[-1] != '=' && p[-1] != '!' && p[-1] != '<' && p[-1] != '>'))) assigns = true;
	return assigns ? H_CUSTOM : H_NONE;
}

CloRng* rng_new_impl(const char* type, CloRngSeedType seed_type, void* seeds, size_t seeds_count,
		cl_ulong gid_offset, cl_ulong main_seed, const char* hash, CCLContext* ctx, CCLQueue* cq, GError** err) {
	if (err && *err) return NULL;
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	int id = -1;
	for (int i = 0; clo_rng_infos[i].name != NULL; ++i)
		if (type && strcmp(type, clo_rng_infos[i].name) == 0) id = i;
	if (id < 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_IMPL_NOT_FOUND,
			"The requested RNG implementation, '%s', was not found.", type ? type : "(null)");
		return NULL;
	}
	const size_t seed_size = clo_rng_infos[id].seed_size;
	const size_t bytes = seeds_count * seed_size;
	CCLBuffer* dev_seeds = NULL;
	CCLQueue* own_queue = NULL;
	GError* ierr = NULL;
	if (!cq && seed_type != CLO_RNG_SEED_EXT_DEV) {
		own_queue = ccl_queue_new(ctx, NULL, 0, &ierr);
		if (ierr) { g_propagate_error(err, ierr); return NULL; }
		cq = own_queue;
	}
	switch (seed_type) {
	case CLO_RNG_SEED_DEV_GID: {
		if (seeds != NULL) {
			g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "The DEV_GID seed type expects a NULL seeds parameter.");
			break;
		}
		int h = parse_hash(hash, &ierr);
		if (ierr) break;
		dev_seeds = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, bytes, NULL, &ierr);
		if (ierr) break;
		CloDeviceGuard g(ctx->dev.ordinal);
		if (h == H_CUSTOM) {
			void* tmp = NULL;
			std::string msg;
			if (clo_cuda_failed(cudaMallocAsync(&tmp, seeds_count ? seeds_count * 8 : 8, cq->stream), &ierr, "cudaMallocAsync")) break;
			cudaError_t rc = clo_jit_seed_hash(hash, (unsigned long long*) tmp, seeds_count, gid_offset, main_seed, cq->stream, msg);
			if (rc != cudaSuccess) g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg.c_str());
			else clo_cuda_failed(kInitSeeds[id](dev_seeds->ptr, seeds_count, (const u64*) tmp, cq->stream), &ierr, "clo_rng_init");
			cudaFreeAsync(tmp, cq->stream);
			break;
		}
		clo_cuda_failed(kInit[id](dev_seeds->ptr, seeds_count, gid_offset, main_seed, h, cq->stream), &ierr, "clo_rng_init");
		break; }
	case CLO_RNG_SEED_HOST_MT: {
		if (seeds != NULL) {
			g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "The HOST_MT seed type expects a NULL seeds parameter.");
			break;
		}
		dev_seeds = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, bytes, NULL, &ierr);
		if (ierr) break;
		std::vector<uint32_t> host(bytes / 4);
		HostMT mt((uint32_t) main_seed);
		for (size_t i = 0; i < host.size(); ++i) host[i] = mt.next();
		ccl_buffer_enqueue_write(dev_seeds, cq, CL_TRUE, 0, host.size() * 4, host.data(), NULL, &ierr);
		break; }
	case CLO_RNG_SEED_EXT_DEV: {
		CCLBuffer* ext = (CCLBuffer*) seeds;
		if (!ext || !clo_handle_alive(ext)) {
			g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "The EXT_DEV seed type expects a device buffer.");
			break;
		}
		if (ext->size < bytes) {
			g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS,
				"The '%s' RNG type requires a buffer of at least %d bytes. The size of the "
				"proviced external device seeds buffer is only %d bytes.", type, (int) bytes, (int) ext->size);
			break;
		}
		ccl_buffer_ref(ext);
		dev_seeds = ext;
		break; }
	case CLO_RNG_SEED_EXT_HOST: {
		if (seeds == NULL) {
			g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "The EXT_HOST seed type expects a non-NULL seeds parameter.");
			break;
		}
		dev_seeds = ccl_buffer_new(ctx, CL_MEM_READ_WRITE, bytes, NULL, &ierr);
		if (ierr) break;
		ccl_buffer_enqueue_write(dev_seeds, cq, CL_TRUE, 0, bytes, seeds, NULL, &ierr);
		break; }
	default:
		g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "Unknown seed type.");
	}
	if (!ierr && own_queue) ccl_queue_finish(own_queue, &ierr);
	if (own_queue) ccl_queue_destroy(own_queue);
	if (ierr) {
		if (dev_seeds) ccl_buffer_destroy(dev_seeds);
		g_propagate_error(err, ierr);
		return NULL;
	}
	clo_rng* rng = new clo_rng();
	rng->id = id;
	rng->src = clo_rng_infos[id].src;
	rng->seeds_device = dev_seeds;
	rng->size_in_device = bytes;
	rng->seeds_count = seeds_count;
	rng->ctx = ctx; ccl_context_ref(ctx);
	clo_handle_add(rng);
	return rng;
}
} // namespace

extern "C" CloRng* clo_rng_new(const char* type, CloRngSeedType seed_type, void* seeds, size_t seeds_count,
		cl_ulong main_seed, const char* hash, CCLContext* ctx, CCLQueue* cq, GError** err) {
	return rng_new_impl(type, seed_type, seeds, seeds_count, 0, main_seed, hash, ctx, cq, err);
}

extern "C" CloRng* clo_rng_new_dev_gid_offset(const char* type, size_t seeds_count, cl_ulong gid_offset,
		cl_ulong main_seed, const char* hash, CCLContext* ctx, CCLQueue* cq, GError** err) {
	return rng_new_impl(type, CLO_RNG_SEED_DEV_GID, NULL, seeds_count, gid_offset, main_seed, hash, ctx, cq, err);
}

extern "C" void clo_rng_destroy(CloRng* rng) {
	if (!rng || !clo_handle_remove(rng)) return;
	ccl_buffer_destroy(rng->seeds_device);
	ccl_context_unref(rng->ctx);
	delete rng;
}

extern "C" const char* clo_rng_get_source(CloRng* rng) { return rng ? rng->src : NULL; }
extern "C" CCLBuffer* clo_rng_get_device_seeds(CloRng* rng) { return rng ? rng->seeds_device : NULL; }
extern "C" size_t clo_rng_get_size(CloRng* rng) { return rng ? rng->size_in_device : 0; }

extern "C" CCLEvent* clo_rng_generate(CloRng* rng, CCLQueue* cq, CCLBuffer* out, size_t runs,
		cl_uint bits, cl_uint maxint, GError** err) {
	if (err && *err) return NULL;
	if (!rng || !cq || !out) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "rng generate: NULL argument"); return NULL; }
	if (maxint == 0 && (bits < 1 || bits > 32)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Number of bits must be between 1 and 32.");
		return NULL;
	}
	if (out->size < runs * rng->seeds_count * sizeof(cl_uint)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "rng generate: output buffer too small");
		return NULL;
	}
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq, "clo_rng_generate");
	cudaError_t rc = kGen[rng->id](rng->seeds_device->ptr, rng->seeds_count, runs, out->ptr,
		maxint ? 32u : bits, maxint, cq->stream);
	clo_queue_end(cq, evt);
	if (clo_cuda_failed(rc, err, "clo_rng_generate launch")) return NULL;
	return evt;
}

extern "C" cl_bool clo_rng_generate_host(CloRng* rng, CCLQueue* cq, void* out, size_t runs,
		cl_uint bits, cl_uint maxint, GError** err) {
	if (err && *err) return CL_FALSE;
	if (!rng || !out) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "rng generate: NULL argument"); return CL_FALSE; }
	GError* ierr = NULL;
	CCLQueue* own_queue = NULL;
	CCLBuffer* dev = NULL;
	cl_bool ok = CL_FALSE;
	const size_t bytes = runs * rng->seeds_count * sizeof(cl_uint);
	if (!cq) {
		own_queue = ccl_queue_new(rng->ctx, NULL, 0, &ierr);
		if (ierr) goto done;
		cq = own_queue;
	}
	dev = ccl_buffer_new(rng->ctx, CL_MEM_WRITE_ONLY, bytes, NULL, &ierr);
	if (ierr) goto done;
	clo_rng_generate(rng, cq, dev, runs, bits, maxint, &ierr);
	if (ierr) goto done;
	ccl_buffer_enqueue_read(dev, cq, CL_TRUE, 0, bytes, out, NULL, &ierr);
	if (ierr) goto done;
	ok = CL_TRUE;
done:
	if (ierr) g_propagate_error(err, ierr);
	if (dev) ccl_buffer_destroy(dev);
	if (own_queue) ccl_queue_destroy(own_queue);
	return ok;
}
