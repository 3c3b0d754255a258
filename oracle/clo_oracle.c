/*
 * oracle/clo_oracle.c -- CPU restatement of the cl_ops hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see clo_oracle.h).  Plain C99, no GPU, no product code.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Pinned against the reference's own kernels run on the CPU
 * (oracle/_ref) by tests/golden/make_golden.py -> tests/golden/ref_*.npz.
 */
#include "clo_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* types: src/cl_ops/common/clo_common.c:54-68                                */
/* ------------------------------------------------------------------------- */

static const size_t orc_sizes[11] = { 1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8 };

size_t orc_type_sizeof(int type) {
	return (type >= 0 && type <= 10) ? orc_sizes[type] : 0;
}

static int orc_is_int(int t) { return t >= ORC_CHAR && t <= ORC_ULONG; }
static int orc_is_signed(int t) {
	return t == ORC_CHAR || t == ORC_SHORT || t == ORC_INT || t == ORC_LONG;
}

/* ------------------------------------------------------------------------- */
/* GLib GRand == MT19937 (third-party: GLib >= 2.32.1, CMakeLists.txt:33).    */
/* Restated from glib/grand.c: seeding "version 2.2", genrand_int32 tempering,*/
/* g_rand_double = two draws, g_rand_int_range rejection, g_rand_boolean.     */
/* ------------------------------------------------------------------------- */

#define MT_N 624
#define MT_M 397

void orc_grand_seed(orc_grand* r, uint32_t seed) {
	r->mt[0] = seed;
	for (r->mti = 1; r->mti < MT_N; r->mti++)
		r->mt[r->mti] = 1812433253u *
			(r->mt[r->mti - 1] ^ (r->mt[r->mti - 1] >> 30)) + r->mti;
}

uint32_t orc_grand_int(orc_grand* r) {
	static const uint32_t mag01[2] = { 0x0u, 0x9908b0dfu };
	uint32_t y;
	if (r->mti >= MT_N) {
		int kk;
		for (kk = 0; kk < MT_N - MT_M; kk++) {
			y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
			r->mt[kk] = r->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1];
		}
		for (; kk < MT_N - 1; kk++) {
			y = (r->mt[kk] & 0x80000000u) | (r->mt[kk + 1] & 0x7fffffffu);
			r->mt[kk] = r->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1];
		}
		y = (r->mt[MT_N - 1] & 0x80000000u) | (r->mt[0] & 0x7fffffffu);
		r->mt[MT_N - 1] = r->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1];
		r->mti = 0;
	}
	y = r->mt[r->mti++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	return y;
}

/* G_RAND_DOUBLE_TRANSFORM = 2^-32 */
#define ORC_DT 2.3283064365386962890625e-10

double orc_grand_double(orc_grand* r) {
	double v = orc_grand_int(r) * ORC_DT;
	v = (v + orc_grand_int(r)) * ORC_DT;
	if (v >= 1.0) return orc_grand_double(r);
	return v;
}

int32_t orc_grand_int_range(orc_grand* r, int32_t begin, int32_t end) {
	uint32_t dist = (uint32_t) end - (uint32_t) begin;
	uint32_t random = 0;
	if (dist != 0) {
		uint32_t maxvalue;
		if (dist <= 0x80000000u) {
			/* maxvalue = 2^32 - 1 - (2^32 % dist) */
			uint32_t leftover = (0x80000000u % dist) * 2;
			if (leftover >= dist) leftover -= dist;
			maxvalue = 0xffffffffu - leftover;
		} else {
			maxvalue = dist - 1;
		}
		do random = orc_grand_int(r); while (random > maxvalue);
		random %= dist;
	}
	return (int32_t) ((uint32_t) begin + random);
}

double orc_grand_double_range(orc_grand* r, double begin, double end) {
	double v = orc_grand_double(r);
	return v * end - (v - 1) * begin;
}

int orc_grand_boolean(orc_grand* r) {
	return (orc_grand_int(r) & (1u << 15)) != 0;
}

/* src/benchmarks/clo_bench.c:67-142.  half is not restated (cl_half is an
 * integer typedef in the host headers; the bench never uses it by default). */
void orc_bench_rand(orc_grand* r, int type, void* location) {
	switch (type) {
	case ORC_CHAR: { int8_t v = (int8_t) orc_grand_int_range(r, -128, 127); memcpy(location, &v, 1); break; }
	case ORC_UCHAR: { uint8_t v = (uint8_t) orc_grand_int_range(r, 0, 255); memcpy(location, &v, 1); break; }
	case ORC_SHORT: { int16_t v = (int16_t) orc_grand_int_range(r, -32768, 32767); memcpy(location, &v, 2); break; }
	case ORC_USHORT: { uint16_t v = (uint16_t) orc_grand_int_range(r, 0, 65535); memcpy(location, &v, 2); break; }
	case ORC_INT: { int32_t v = orc_grand_int_range(r, INT32_MIN, INT32_MAX); memcpy(location, &v, 4); break; }
	case ORC_UINT: { uint32_t v = (uint32_t) (orc_grand_double(r) * 0xffffffffu); memcpy(location, &v, 4); break; }
	case ORC_LONG: {
		/* C evaluation order of the two GRand calls in
		 * g_rand_double(rng) * (g_rand_boolean(rng) ? MIN : MAX) is unspecified;
		 * gcc evaluates the left operand first. */
		double d = orc_grand_double(r);
		int b = orc_grand_boolean(r);
		int64_t v = (int64_t) (d * (b ? (double) INT64_MIN : (double) INT64_MAX));
		memcpy(location, &v, 8); break; }
	case ORC_ULONG: {
		double d = orc_grand_double(r) * (double) UINT64_MAX;
		uint64_t v = (uint64_t) d; memcpy(location, &v, 8); break; }
	case ORC_FLOAT: { float v = (float) orc_grand_double_range(r, 1.175494350822287507969e-38, 340282346638528859811704183484516925440.0); memcpy(location, &v, 4); break; }
	case ORC_DOUBLE: { double v = orc_grand_double_range(r, 2.225073858507201383090e-308, 179769313486231570814527423731704356798070567525844996598917476803157260780028538760589558632766878171540458953514382464234321326889464182768467546703537516986049910576551282076245490090389328944075868508455133942304583236903222948165808559332123348274797826204144723168738177180919299881250404026184124858368.0); memcpy(location, &v, 8); break; }
	default: break;
	}
}

void orc_fill_sort_input(uint32_t seed, int type, void* data, size_t n) {
	orc_grand r; size_t sz = orc_type_sizeof(type);
	orc_grand_seed(&r, seed);
	for (size_t i = 0; i < n; i++) orc_bench_rand(&r, type, (char*) data + i * sz);
}

/* src/benchmarks/clo_scan_bench.c:219-224: gulong value = g_rand_double*128,
 * memcpy of the low `bytes` bytes. */
void orc_fill_scan_input(uint32_t seed, int type, void* data, size_t n) {
	orc_grand r; size_t sz = orc_type_sizeof(type);
	orc_grand_seed(&r, seed);
	for (size_t i = 0; i < n; i++) {
		uint64_t value = (uint64_t) (orc_grand_double(&r) * 128);
		memcpy((char*) data + i * sz, &value, sz);
	}
}

/* ------------------------------------------------------------------------- */
/* RNG                                                                        */
/* ------------------------------------------------------------------------- */

/* seed sizes: src/cl_ops/rng/clo_rng.c:60-68 */
static const size_t orc_seed_sizes[6] = { 8, 8, 16, 8, 4, 16 };

size_t orc_rng_seed_size(int rng) {
	return (rng >= 0 && rng < 6) ? orc_seed_sizes[rng] : 0;
}

/* src/cl_ops/rng/clo_rng_init.cl:29-41 -- all arithmetic on a 64-bit ulong. */
static uint64_t orc_hash(int hash, uint64_t x) {
	switch (hash) {
	case ORC_HASH_KNUTH:
		x = ((x * 2654435761ull) % 0x100000000ull);
		break;
	case ORC_HASH_XS1:
		x = ((x >> 16) ^ x) * 0x45d9f3bull;
		x = ((x >> 16) ^ x) * 0x45d9f3bull;
		x = ((x >> 16) ^ x);
		break;
	default: break;
	}
	return x;
}

/* clo_ulong2statetype: lcg.cl:32, xorshift64.cl:29, xorshift128.cl:30 (note >> 46),
 * mwc64x.cl:32 (as_uint2, little endian), parkmiller.cl:31, tauslcg.cl:35. */
static void orc_ulong2state(int rng, uint64_t seed, void* state) {
	switch (rng) {
	case ORC_RNG_LCG:
	case ORC_RNG_XORSHIFT64:
		memcpy(state, &seed, 8); break;
	case ORC_RNG_XORSHIFT128: {
		uint32_t s[4];
		s[0] = (uint32_t) (0xFFFFFFFFull & seed);
		s[1] = (uint32_t) (0xFFFFFFFFull & (seed >> 16));
		s[2] = (uint32_t) (0xFFFFFFFFull & (seed >> 32));
		s[3] = (uint32_t) (0xFFFFFFFFull & (seed >> 46));
		memcpy(state, s, 16); break; }
	case ORC_RNG_MWC64X: {
		uint32_t s[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
		memcpy(state, s, 8); break; }
	case ORC_RNG_PARKMILLER: {
		uint32_t s = (uint32_t) (0xFFFFFFFFull & seed);
		memcpy(state, &s, 4); break; }
	case ORC_RNG_TAUSLCG: {
		uint32_t s[4] = { (uint32_t) seed, (uint32_t) (seed >> 32),
			(uint32_t) seed, (uint32_t) (seed >> 32) };
		memcpy(state, s, 16); break; }
	default: break;
	}
}

/* src/cl_ops/rng/clo_rng_init.cl:47-60 */
void orc_rng_seed_dev_gid(int rng, int hash, uint64_t main_seed,
		uint64_t gid0, size_t count, void* states) {
	size_t ss = orc_rng_seed_size(rng);
	for (size_t i = 0; i < count; i++) {
		uint64_t seed = (gid0 + i) + main_seed;
		seed = orc_hash(hash, seed);
		orc_ulong2state(rng, seed, (char*) states + i * ss);
	}
}

/* src/cl_ops/rng/clo_rng.c:185-203: g_rand_new_with_seed((guint32) main_seed),
 * one g_rand_int per 4 bytes of the seeds vector, memory order. */
void orc_rng_seed_host_mt(int rng, uint64_t main_seed, size_t count, void* states) {
	orc_grand r; size_t words = orc_rng_seed_size(rng) * count / 4;
	orc_grand_seed(&r, (uint32_t) main_seed);
	for (size_t i = 0; i < words; i++) {
		uint32_t v = orc_grand_int(&r);
		memcpy((char*) states + i * 4, &v, 4);
	}
}

/* tauslcg.cl:52-55 */
static uint32_t orc_taus_step(uint32_t z, int s1, int s2, int s3, uint32_t m) {
	uint32_t b = (((z << s1) ^ z) >> s2);
	return (((z & m) << s3) ^ b);
}

uint32_t orc_rng_next(int rng, void* state) {
	switch (rng) {
	case ORC_RNG_LCG: { /* clo_rng_lcg.cl:43-59 */
		uint64_t s; memcpy(&s, state, 8);
		s = (s * 0x5DEECE66Dull + 0xBull) & ((1ull << 48) - 1);
		memcpy(state, &s, 8);
		return (uint32_t) (s >> (48 - 32)); }
	case ORC_RNG_XORSHIFT64: { /* clo_rng_xorshift64.cl:40-63 */
		uint64_t s; memcpy(&s, state, 8);
		s ^= (s << 21); s ^= (s >> 35); s ^= (s << 4);
		memcpy(state, &s, 8);
		return (uint32_t) s; }
	case ORC_RNG_XORSHIFT128: { /* clo_rng_xorshift128.cl:41-59 */
		uint32_t s[4]; memcpy(s, state, 16);
		uint32_t t = s[0] ^ (s[0] << 11);
		s[0] = s[1]; s[1] = s[2]; s[2] = s[3];
		s[3] = s[3] ^ (s[3] >> 19) ^ (t ^ (t >> 8));
		memcpy(state, s, 16);
		return s[3]; }
	case ORC_RNG_MWC64X: { /* clo_rng_mwc64x.cl:43-63 */
		uint32_t s[2]; memcpy(s, state, 8);
		const uint32_t A = 4294883355u;
		uint32_t x = s[0], c = s[1];
		uint32_t res = x ^ c;
		uint32_t hi = (uint32_t) (((uint64_t) x * A) >> 32);
		x = x * A + c;
		c = hi + (x < c);
		s[0] = x; s[1] = c;
		memcpy(state, s, 8);
		return res; }
	case ORC_RNG_PARKMILLER: { /* clo_rng_parkmiller.cl:43-59: C signed remainder */
		int32_t s; memcpy(&s, state, 4);
		s = (int32_t) ((((int64_t) s) * 16807) % 2147483647);
		memcpy(state, &s, 4);
		return ((uint32_t) s) << 1; }
	case ORC_RNG_TAUSLCG: { /* clo_rng_tauslcg.cl:79-100 */
		uint32_t s[4]; memcpy(s, state, 16);
		uint32_t x = s[0];
		s[0] = orc_taus_step(s[1], 13, 19, 12, 4294967294u);
		s[1] = orc_taus_step(s[2], 2, 25, 4, 4294967288u);
		s[2] = orc_taus_step(s[3], 3, 11, 17, 4294967294u);
		s[3] = 1664525u * x + 1013904223u;
		memcpy(state, s, 16);
		return s[0]; }
	default: return 0;
	}
}

/* src/benchmarks/clo_rng_bench.cl:23-37 (one launch = one run over all streams),
 * src/benchmarks/clo_rng_bench.c:302-324 (runs loop; row r of the output),
 * clo_rng_api.cl:33-40 (next_int = next % n). */
void orc_rng_generate(int rng, void* states, size_t count, size_t runs,
		uint32_t bits, uint32_t maxint, uint32_t* out) {
	size_t ss = orc_rng_seed_size(rng);
	for (size_t r = 0; r < runs; r++) {
		for (size_t g = 0; g < count; g++) {
			uint32_t v = orc_rng_next(rng, (char*) states + g * ss);
			out[r * count + g] = maxint ? (v % maxint) : (v >> (32 - bits));
		}
	}
}

/* ------------------------------------------------------------------------- */
/* scan                                                                       */
/* ------------------------------------------------------------------------- */

/* integer load, widened to 64 bits the way the implicit elem->sum conversion
 * in clo_scan_blelloch.cl:79-80 does (sign-extend signed, zero-extend unsigned) */
static uint64_t orc_load_int(int t, const void* p, size_t i) {
	switch (t) {
	case ORC_CHAR: return (uint64_t) (int64_t) ((const int8_t*) p)[i];
	case ORC_UCHAR: return ((const uint8_t*) p)[i];
	case ORC_SHORT: return (uint64_t) (int64_t) ((const int16_t*) p)[i];
	case ORC_USHORT: return ((const uint16_t*) p)[i];
	case ORC_INT: return (uint64_t) (int64_t) ((const int32_t*) p)[i];
	case ORC_UINT: return ((const uint32_t*) p)[i];
	case ORC_LONG: return (uint64_t) ((const int64_t*) p)[i];
	case ORC_ULONG: return ((const uint64_t*) p)[i];
	default: return 0;
	}
}

static void orc_store_int(int t, void* p, size_t i, uint64_t v) {
	switch (orc_type_sizeof(t)) {
	case 1: ((uint8_t*) p)[i] = (uint8_t) v; break;
	case 2: ((uint16_t*) p)[i] = (uint16_t) v; break;
	case 4: ((uint32_t*) p)[i] = (uint32_t) v; break;
	case 8: ((uint64_t*) p)[i] = v; break;
	default: break;
	}
}

static double orc_load_num(int t, const void* p, size_t i) {
	if (t == ORC_FLOAT) return ((const float*) p)[i];
	if (t == ORC_DOUBLE) return ((const double*) p)[i];
	if (orc_is_signed(t)) return (double) (int64_t) orc_load_int(t, p, i);
	return (double) orc_load_int(t, p, i);
}

/* Exclusive prefix sum in SUM-type arithmetic.  Integer sums wrap modulo
 * 2^(8*sizeof(sum)) -- truncation commutes with addition, so a 64-bit running
 * sum truncated on store is exactly what the tree in clo_scan_blelloch.cl:82-117
 * produces for any tiling.  Float sums here are the SERIAL association order
 * (the reference's tree order depends on lws; float parity is by tolerance). */
int orc_scan(int elem_type, int sum_type, const void* in, void* out, size_t n) {
	if (orc_is_int(sum_type) && orc_is_int(elem_type)) {
		uint64_t acc = 0;
		for (size_t i = 0; i < n; i++) {
			orc_store_int(sum_type, out, i, acc);
			acc += orc_load_int(elem_type, in, i);
		}
		return 0;
	}
	if (sum_type == ORC_FLOAT) {
		float acc = 0;
		for (size_t i = 0; i < n; i++) {
			((float*) out)[i] = acc;
			acc += (float) orc_load_num(elem_type, in, i);
		}
		return 0;
	}
	if (sum_type == ORC_DOUBLE) {
		double acc = 0;
		for (size_t i = 0; i < n; i++) {
			((double*) out)[i] = acc;
			acc += orc_load_num(elem_type, in, i);
		}
		return 0;
	}
	return -1;
}

void orc_scan_f64ref(int elem_type, const void* in, double* out, size_t n) {
	double acc = 0;
	for (size_t i = 0; i < n; i++) { out[i] = acc; acc += orc_load_num(elem_type, in, i); }
}

/* Three-phase Blelloch restatement: clo_scan_blelloch.c:130-141 (sizing),
 * clo_scan_blelloch.cl:49-126 (workgroupScan: each WG walks blocks_per_wg blocks
 * of 2*lws elements with the up-sweep/down-sweep tree and a running in_sum),
 * :134-182 (scan of WG sums), :193-211 (add WG sums).  The tree is executed
 * literally so float results carry the reference's association order. */
#define ORC_BLELLOCH_BODY(T)                                                         \
	size_t block = 2 * lws;                                                          \
	size_t realws = n / 2;                                                           \
	size_t gws1 = ((realws + lws - 1) / lws) * lws;                                  \
	if (gws1 > lws * lws) gws1 = lws * lws;                                          \
	size_t nwg = gws1 / lws;                                                         \
	size_t bpw = (realws + gws1 - 1) / gws1;                                         \
	size_t nblocks = n / block;                                                      \
	T* wgsums = (T*) calloc(nwg > 2 * lws ? nwg : 2 * lws, sizeof(T));               \
	if (!wgsums) return -1;                                                          \
	_Pragma("omp parallel for schedule(static) num_threads(threads)")                \
	for (long wg = 0; wg < (long) nwg; wg++) {                                       \
		T* aux = (T*) malloc(block * sizeof(T));                                     \
		T in_sum = 0;                                                                \
		for (size_t b = 0; b < bpw && (wg * bpw + b) < nblocks; b++) {               \
			size_t g0 = (bpw * wg + b) * block;                                      \
			for (size_t k = 0; k < block; k++) aux[k] = src[g0 + k];                 \
			size_t offset = 1;                                                       \
			for (size_t d = block >> 1; d > 0; d >>= 1) {                            \
				for (size_t lid = 0; lid < d; lid++)                                 \
					aux[offset * (2 * lid + 2) - 1] += aux[offset * (2 * lid + 1) - 1]; \
				offset *= 2;                                                         \
			}                                                                        \
			T prev = in_sum;                                                         \
			in_sum += aux[block - 1];                                                \
			aux[block - 1] = 0;                                                      \
			for (size_t d = 1; d < block; d *= 2) {                                  \
				offset >>= 1;                                                        \
				for (size_t lid = 0; lid < d; lid++) {                               \
					size_t ai = offset * (2 * lid + 1) - 1;                          \
					size_t bi = offset * (2 * lid + 2) - 1;                          \
					T t = aux[ai]; aux[ai] = aux[bi]; aux[bi] += t;                  \
				}                                                                    \
			}                                                                        \
			for (size_t k = 0; k < block; k++) dst[g0 + k] = aux[k] + prev;          \
		}                                                                            \
		wgsums[wg] = in_sum;                                                         \
		free(aux);                                                                   \
	}                                                                                \
	if (gws1 > lws) {                                                                \
		/* workgroupSumsScan: one WG of (nwg/2) items over nwg sums (tree) */        \
		size_t blk2 = nwg;                                                           \
		size_t offset = 1;                                                           \
		for (size_t d = blk2 >> 1; d > 0; d >>= 1) {                                 \
			for (size_t lid = 0; lid < d; lid++)                                     \
				wgsums[offset * (2 * lid + 2) - 1] += wgsums[offset * (2 * lid + 1) - 1]; \
			offset *= 2;                                                             \
		}                                                                            \
		wgsums[blk2 - 1] = 0;                                                        \
		for (size_t d = 1; d < blk2; d *= 2) {                                       \
			offset >>= 1;                                                            \
			for (size_t lid = 0; lid < d; lid++) {                                   \
				size_t ai = offset * (2 * lid + 1) - 1;                              \
				size_t bi = offset * (2 * lid + 2) - 1;                              \
				T t = wgsums[ai]; wgsums[ai] = wgsums[bi]; wgsums[bi] += t;          \
			}                                                                        \
		}                                                                            \
		/* addWorkgroupSums: group g of lws items adds wgsums[g / (2*bpw)] */        \
		_Pragma("omp parallel for schedule(static) num_threads(threads)")            \
		for (long i = 0; i < (long) n; i++)                                          \
			dst[i] += wgsums[((size_t) i / lws) / (2 * bpw)];                        \
	}                                                                                \
	free(wgsums);                                                                    \
	return 0;

static int orc_blelloch_u32(const uint32_t* src, uint32_t* dst, size_t n, size_t lws, int threads) {
	ORC_BLELLOCH_BODY(uint32_t)
}
static int orc_blelloch_f32(const float* src, float* dst, size_t n, size_t lws, int threads) {
	ORC_BLELLOCH_BODY(float)
}

int orc_scan_blelloch_port(int elem_type, const void* in, void* out, size_t n,
		size_t lws, int threads) {
	/* parity domain: power-of-two n >= 2*lws */
	if (lws == 0 || (lws & (lws - 1)) || n < 2 * lws || (n & (n - 1))) return -1;
	if (threads < 1) threads = 1;
	if (elem_type == ORC_UINT) return orc_blelloch_u32((const uint32_t*) in, (uint32_t*) out, n, lws, threads);
	if (elem_type == ORC_FLOAT) return orc_blelloch_f32((const float*) in, (float*) out, n, lws, threads);
	return -1;
}

/* ------------------------------------------------------------------------- */
/* sort                                                                       */
/* ------------------------------------------------------------------------- */

static uint64_t orc_load_raw(size_t sz, const void* p, size_t i) {
	switch (sz) {
	case 1: return ((const uint8_t*) p)[i];
	case 2: return ((const uint16_t*) p)[i];
	case 4: return ((const uint32_t*) p)[i];
	default: return ((const uint64_t*) p)[i];
	}
}

/* CLO_SORT_KEY_TYPE key = CLO_SORT_KEY_GET(elem)  (clo_sort_sbitonic.cl:53-54).
 * Returns the key's raw bits, zero-extended, truncated to the key type. */
static uint64_t orc_key(const orc_sortspec* s, const void* data, size_t i) {
	size_t esz = orc_type_sizeof(s->elem_type), ksz = orc_type_sizeof(s->key_type);
	uint64_t k;
	if (!orc_is_int(s->elem_type)) {
		/* float/double element: identity key only */
		return orc_load_raw(esz, data, i);
	}
	uint64_t x = orc_load_int(s->elem_type, data, i);
	if (orc_is_signed(s->elem_type)) k = (uint64_t) (((int64_t) x) >> s->shift);
	else k = x >> s->shift;
	k &= s->mask;
	if (ksz < 8) k &= ((1ull << (8 * ksz)) - 1);
	return k;
}

/* CLO_SORT_COMPARE(a, b): default "((a) > (b))" (clo_sort_abstract.c:157-161);
 * descending menu entry "((a) < (b))".  Typed comparison on KEY_TYPE. */
static int orc_cmp(const orc_sortspec* s, uint64_t a, uint64_t b) {
	int gt, lt;
	switch (s->key_type) {
	case ORC_CHAR: gt = (int8_t) a > (int8_t) b; lt = (int8_t) a < (int8_t) b; break;
	case ORC_SHORT: gt = (int16_t) a > (int16_t) b; lt = (int16_t) a < (int16_t) b; break;
	case ORC_INT: gt = (int32_t) a > (int32_t) b; lt = (int32_t) a < (int32_t) b; break;
	case ORC_LONG: gt = (int64_t) a > (int64_t) b; lt = (int64_t) a < (int64_t) b; break;
	case ORC_FLOAT: { float fa, fb; uint32_t ua = (uint32_t) a, ub = (uint32_t) b;
		memcpy(&fa, &ua, 4); memcpy(&fb, &ub, 4); gt = fa > fb; lt = fa < fb; break; }
	case ORC_DOUBLE: { double fa, fb; memcpy(&fa, &a, 8); memcpy(&fb, &b, 8);
		gt = fa > fb; lt = fa < fb; break; }
	default: gt = a > b; lt = a < b; break;
	}
	return s->descending ? lt : gt;
}

static void orc_swap(size_t sz, void* data, size_t i, size_t j) {
	unsigned char t[8];
	memcpy(t, (char*) data + i * sz, sz);
	memcpy((char*) data + i * sz, (char*) data + j * sz, sz);
	memcpy((char*) data + j * sz, t, sz);
}

static size_t orc_nlpo2(size_t x) { size_t p = 1; while (p < x) p <<= 1; return p; }

/* Canonical bitonic network.  Kernel: clo_sort_sbitonic.cl:38-69; host loop:
 * clo_sort_sbitonic.c:73-118 (gws = nlpo2(n)/2, stages 1..log2(2*gws), steps
 * stage..1).  For n not a power of two the reference indexes past the buffer;
 * here out-of-range partners are treated as +infinity sentinels (never swapped
 * into range), which is what any correct ascending sort must equal. */
int orc_sort_bitonic(const orc_sortspec* s, void* data, size_t n) {
	size_t esz = orc_type_sizeof(s->elem_type);
	size_t np2 = orc_nlpo2(n), gws = np2 / 2;
	unsigned tot_stages = 0;
	if (!esz) return -1;
	while (((size_t) 1 << tot_stages) < np2) tot_stages++;
	void* buf = data; size_t* idx = NULL;
	if (np2 != n) {
		/* pad with sentinels flagged by index >= n */
		buf = malloc(np2 * esz);
		if (!buf) return -1;
		memcpy(buf, data, n * esz);
		idx = (size_t*) malloc(np2 * sizeof(size_t));
		if (!idx) { free(buf); return -1; }
		for (size_t i = 0; i < np2; i++) idx[i] = i;
	}
	for (unsigned stage = 1; stage <= tot_stages; stage++) {
		for (unsigned step = stage; step > 0; step--) {
			size_t pair_stride = (size_t) 1 << (step - 1);
			for (size_t gid = 0; gid < gws; gid++) {
				size_t index1 = gid + (gid / pair_stride) * pair_stride;
				size_t index2 = index1 + pair_stride;
				int desc = (int) (1 & (gid >> (stage - 1)));
				int cmp;
				if (idx) {
					int pad1 = idx[index1] >= n, pad2 = idx[index2] >= n;
					if (pad1 || pad2) cmp = pad1 && !pad2; /* sentinel is larger */
					else cmp = orc_cmp(s, orc_key(s, buf, index1), orc_key(s, buf, index2));
				} else {
					cmp = orc_cmp(s, orc_key(s, buf, index1), orc_key(s, buf, index2));
				}
				if (cmp ^ desc) {
					orc_swap(esz, buf, index1, index2);
					if (idx) { size_t t = idx[index1]; idx[index1] = idx[index2]; idx[index2] = t; }
				}
			}
		}
	}
	if (idx) { memcpy(data, buf, n * esz); free(buf); free(idx); }
	return 0;
}

/* clo_sort_gselect.cl:38-57: rank = #{i : COMPARE(key_gid, key_i) ||
 * (key_i == key_gid && i < gid)}; out[rank] = in[gid]. */
int orc_sort_gselect(const orc_sortspec* s, const void* in, void* out, size_t n) {
	size_t esz = orc_type_sizeof(s->elem_type);
	if (!esz) return -1;
	for (size_t gid = 0; gid < n; gid++) {
		size_t pos = 0;
		uint64_t kg = orc_key(s, in, gid);
		for (size_t i = 0; i < n; i++) {
			uint64_t ki = orc_key(s, in, i);
			if (orc_cmp(s, kg, ki) || ((ki == kg) && (i < gid))) pos++;
		}
		memcpy((char*) out + pos * esz, (const char*) in + gid * esz, esz);
	}
	return 0;
}

/* bit b of the key as the kernels read it: (key >> b) & 1 with OpenCL C shift
 * semantics (count taken modulo the promoted width; arithmetic for signed keys):
 * clo_sort_satradix.cl:58-61,143-144,248-249. */
static unsigned orc_digit(const orc_sortspec* s, uint64_t key, unsigned start_bit, unsigned nbits) {
	size_t ksz = orc_type_sizeof(s->key_type);
	unsigned W = ksz == 8 ? 64 : 32;
	uint64_t promoted = key;
	if (orc_is_signed(s->key_type) && ksz < 8) {
		/* sign-extend to the promoted width */
		unsigned kb = 8 * (unsigned) ksz;
		if (key & (1ull << (kb - 1))) promoted |= ~((1ull << kb) - 1);
		if (W == 32) promoted &= 0xffffffffull;
	}
	unsigned sh = start_bit & (W - 1);
	uint64_t v;
	if (orc_is_signed(s->key_type)) {
		if (W == 64) v = (uint64_t) (((int64_t) promoted) >> sh);
		else v = (uint64_t) (uint32_t) (((int32_t) (uint32_t) promoted) >> sh);
	} else v = promoted >> sh;
	return (unsigned) (v & ((1u << nbits) - 1));
}

static unsigned orc_tzc(unsigned x) { unsigned c = 0; while (x && !(x & 1)) { x >>= 1; c++; } return c; }

/* Satish-style LSD radix, tile structure restated:
 *  host loop clo_sort_satradix.c:264-313 (total_digits = elem_bits / bits_in_digit,
 *  :166-169); per pass: localsort (tile-local stable sort by the digit,
 *  clo_sort_satradix.cl:34-123), histogram (offsets[wg][d], counters[d][wg],
 *  :125-222), exclusive scan of counters (clo_sort_satradix.c:298-299), scatter
 *  (out[counters_sum[d][wg] + lid - offsets[wg][d]], :224-258).
 * Keys are raw bits, ascending; CLO_SORT_COMPARE is ignored by the reference. */
int orc_sort_satradix(const orc_sortspec* s, void* data, size_t n,
		uint32_t radix, size_t lws, int threads) {
	size_t esz = orc_type_sizeof(s->elem_type);
	unsigned nbits = orc_tzc(radix);
	if (!esz || radix < 2 || (radix & (radix - 1))) return -1;
	unsigned total_digits = (unsigned) (esz * 8 / nbits);
	if (threads < 1) threads = 1;
	if (n < 2) return 0;
	/* tile emulation on the parity domain; any other n: single tile (results are
	 * tiling independent because every step is stable) */
	if (lws < radix) lws = radix;
	if ((n & (n - 1)) || lws > n || (lws & (lws - 1))) lws = n;
	size_t num_wgs = n / lws;
	void* tmp = malloc(n * esz);
	uint32_t* offsets = (uint32_t*) malloc(num_wgs * radix * sizeof(uint32_t));
	uint64_t* counters = (uint64_t*) malloc(num_wgs * radix * sizeof(uint64_t));
	if (!tmp || !offsets || !counters) { free(tmp); free(offsets); free(counters); return -1; }
	for (unsigned pass = 0; pass < total_digits; pass++) {
		unsigned start_bit = pass * nbits;
		/* localsort + histogram */
		#pragma omp parallel for schedule(static) num_threads(threads)
		for (long wg = 0; wg < (long) num_wgs; wg++) {
			size_t base = (size_t) wg * lws;
			uint32_t* cnt = (uint32_t*) calloc(radix, sizeof(uint32_t));
			uint32_t* off = offsets + (size_t) wg * radix;
			for (size_t l = 0; l < lws; l++)
				cnt[orc_digit(s, orc_key(s, data, base + l), start_bit, nbits)]++;
			uint32_t run = 0;
			for (uint32_t d = 0; d < radix; d++) {
				off[d] = run;
				counters[(size_t) d * num_wgs + wg] = cnt[d];
				run += cnt[d];
				cnt[d] = off[d];
			}
			for (size_t l = 0; l < lws; l++) {
				unsigned d = orc_digit(s, orc_key(s, data, base + l), start_bit, nbits);
				memcpy((char*) tmp + (base + cnt[d]++) * esz, (char*) data + (base + l) * esz, esz);
			}
			free(cnt);
		}
		/* global exclusive scan of the digit-major counters */
		uint64_t acc = 0;
		for (size_t i = 0; i < num_wgs * radix; i++) { uint64_t c = counters[i]; counters[i] = acc; acc += c; }
		/* scatter */
		#pragma omp parallel for schedule(static) num_threads(threads)
		for (long wg = 0; wg < (long) num_wgs; wg++) {
			size_t base = (size_t) wg * lws;
			const uint32_t* off = offsets + (size_t) wg * radix;
			for (size_t l = 0; l < lws; l++) {
				unsigned d = orc_digit(s, orc_key(s, tmp, base + l), start_bit, nbits);
				size_t out_idx = (size_t) counters[(size_t) d * num_wgs + wg] + l - off[d];
				memcpy((char*) data + out_idx * esz, (char*) tmp + (base + l) * esz, esz);
			}
		}
	}
	free(tmp); free(offsets); free(counters);
	return 0;
}

/* Stable LSD (8-bit digits) of separate key / payload arrays, raw unsigned key
 * bits ascending -- the semantics satradix gives a packed (key|payload) element,
 * stated for the additive pairs entry point. */
int orc_sort_pairs(int key_type, void* keys, uint32_t* payload, size_t n) {
	size_t ksz = orc_type_sizeof(key_type);
	if (ksz != 4 && ksz != 8) return -1;
	void* k2 = malloc(n * ksz);
	uint32_t* p2 = (uint32_t*) malloc(n * sizeof(uint32_t));
	size_t* cnt = (size_t*) malloc(256 * sizeof(size_t));
	if (!k2 || !p2 || !cnt) { free(k2); free(p2); free(cnt); return -1; }
	void* ka = keys; void* kb = k2; uint32_t* pa = payload; uint32_t* pb = p2;
	for (unsigned pass = 0; pass < ksz; pass++) {
		unsigned sh = 8 * pass;
		memset(cnt, 0, 256 * sizeof(size_t));
		for (size_t i = 0; i < n; i++) cnt[(orc_load_raw(ksz, ka, i) >> sh) & 255]++;
		size_t run = 0;
		for (unsigned d = 0; d < 256; d++) { size_t c = cnt[d]; cnt[d] = run; run += c; }
		for (size_t i = 0; i < n; i++) {
			uint64_t k = orc_load_raw(ksz, ka, i);
			size_t o = cnt[(k >> sh) & 255]++;
			if (ksz == 4) ((uint32_t*) kb)[o] = (uint32_t) k; else ((uint64_t*) kb)[o] = k;
			pb[o] = pa[i];
		}
		void* t = ka; ka = kb; kb = t;
		uint32_t* tp = pa; pa = pb; pb = tp;
	}
	/* ksz passes is even for 4 and 8 -> result is back in keys/payload */
	free(k2); free(p2); free(cnt);
	return 0;
}
