/*
 * oracle/clo_oracle.h -- CPU oracle for the cl_ops hot path (sort / scan / bulk RNG).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cl_ops_b200/, include/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as a checker
 * or as the timed CPU baseline -- never as the thing shipped.
 *
 * Parity status: the restatement is PINNED against the reference's own OpenCL C
 * kernels executed on the CPU (oracle/_ref, built by oracle/build_ref.sh from the
 * .cl files where they lie under /root/reference) -- see tests/golden/make_golden.py
 * and tests/test_oracle_golden.py.  The GLib GRand (MT19937) part is a third-party
 * dependency absent from /root/reference; it is restated from GLib's published
 * algorithm (grand.c, GLib >= 2.32) and cross-checked against numpy's MT19937.
 */
#ifndef CLO_ORACLE_H
#define CLO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CloType values: src/cl_ops/common/clo_common.in.h:108-120 */
enum {
	ORC_CHAR = 0, ORC_UCHAR = 1, ORC_SHORT = 2, ORC_USHORT = 3, ORC_INT = 4,
	ORC_UINT = 5, ORC_LONG = 6, ORC_ULONG = 7, ORC_HALF = 8, ORC_FLOAT = 9,
	ORC_DOUBLE = 10
};

/* RNG ids, order of clo_rng_infos[]: src/cl_ops/rng/clo_rng.c:60-68 */
enum {
	ORC_RNG_LCG = 0, ORC_RNG_XORSHIFT64 = 1, ORC_RNG_XORSHIFT128 = 2,
	ORC_RNG_MWC64X = 3, ORC_RNG_PARKMILLER = 4, ORC_RNG_TAUSLCG = 5
};

/* Seed hashes: src/cl_ops/rng/clo_rng_init.cl:27-41 */
enum { ORC_HASH_NONE = 0, ORC_HASH_KNUTH = 1, ORC_HASH_XS1 = 2 };

size_t orc_type_sizeof(int type);

/* ---- GLib GRand (MT19937) restatement ---- */
typedef struct orc_grand { uint32_t mt[624]; uint32_t mti; } orc_grand;
void     orc_grand_seed(orc_grand* r, uint32_t seed);
uint32_t orc_grand_int(orc_grand* r);
double   orc_grand_double(orc_grand* r);
int32_t  orc_grand_int_range(orc_grand* r, int32_t begin, int32_t end);
double   orc_grand_double_range(orc_grand* r, double begin, double end);
int      orc_grand_boolean(orc_grand* r);
/* clo_bench_rand: src/benchmarks/clo_bench.c:67-142 (one element) */
void     orc_bench_rand(orc_grand* r, int type, void* location);
/* fills as the sort bench does (clo_sort_bench.c:190-193) / scan bench (clo_scan_bench.c:219-224) */
void     orc_fill_sort_input(uint32_t seed, int type, void* data, size_t n);
void     orc_fill_scan_input(uint32_t seed, int type, void* data, size_t n);

/* ---- RNG ---- */
size_t   orc_rng_seed_size(int rng);
/* clo_rng_init kernel + clo_ulong2statetype: states[gid], gid in [gid0, gid0+count) */
void     orc_rng_seed_dev_gid(int rng, int hash, uint64_t main_seed,
			uint64_t gid0, size_t count, void* states);
/* clo_rng_host_seed_init: src/cl_ops/rng/clo_rng.c:185-203 */
void     orc_rng_seed_host_mt(int rng, uint64_t main_seed, size_t count, void* states);
/* one clo_rng_next on the state at `state` (updates it) */
uint32_t orc_rng_next(int rng, void* state);
/* bulk generation, clo_rng_bench.cl:23-37 + clo_rng_bench.c:302-324:
 * out[r*count + g] = next_r(state_g) >> (32-bits)     (maxint == 0)
 *                  = next_r(state_g) % maxint         (maxint != 0) */
void     orc_rng_generate(int rng, void* states, size_t count, size_t runs,
			uint32_t bits, uint32_t maxint, uint32_t* out);

/* ---- scan ---- */
/* exclusive prefix sum, SUM-type arithmetic: clo_scan_blelloch.cl:66-125,198-209 */
int      orc_scan(int elem_type, int sum_type, const void* in, void* out, size_t n);
/* double-precision reference for float tolerance tests */
void     orc_scan_f64ref(int elem_type, const void* in, double* out, size_t n);
/* three-phase Blelloch restatement (u32->u32 and f32->f32 only), `threads` host threads;
 * used for the float association order and as the timed CPU baseline. */
int      orc_scan_blelloch_port(int elem_type, const void* in, void* out, size_t n,
			size_t lws, int threads);

/* ---- sort ---- */
/* Key extraction menu standing in for the CLO_SORT_KEY_GET macro string
 * (clo_sort_abstract.c:162-168): key = (KEY_TYPE)(((x) >> shift) & mask). */
typedef struct orc_sortspec {
	int elem_type;
	int key_type;
	uint32_t shift;
	uint64_t mask;      /* all ones = no mask */
	int descending;     /* 0: compare "((a) > (b))" (default), 1: "((a) < (b))" */
} orc_sortspec;

/* canonical bitonic network: clo_sort_sbitonic.cl:38-69 + clo_sort_sbitonic.c:73-118 */
int      orc_sort_bitonic(const orc_sortspec* s, void* data, size_t n);
/* stable rank sort: clo_sort_gselect.cl:38-57 */
int      orc_sort_gselect(const orc_sortspec* s, const void* in, void* out, size_t n);
/* stable LSD radix on raw key bits, ascending: clo_sort_satradix.cl:34-258,
 * clo_sort_satradix.c:166-169,264-313 (tile structure: lws, radix) */
int      orc_sort_satradix(const orc_sortspec* s, void* data, size_t n,
			uint32_t radix, size_t lws, int threads);
/* stable sort of (key,payload) pairs held in separate arrays (additive API, C3) */
int      orc_sort_pairs(int key_type, void* keys, uint32_t* payload, size_t n);

#ifdef __cplusplus
}
#endif
#endif
