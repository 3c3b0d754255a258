/*
 * include/cl_ops/clo_rng.h -- the RNG object (device seeds buffer + generator id).
 * Replaces: /root/reference/src/cl_ops/rng/clo_rng.in.h:50-111
 */
#ifndef CLO_B200_RNG_H
#define CLO_B200_RNG_H

#include <cl_ops/clo_common.h>

#ifdef __cplusplus
extern "C" {
#endif

/* clo_rng.in.h:47 */
#define CLO_RNG_IMPLS "lcg, xorshift64, xorshift128, mwc64x, parkmiller, tauslcg"

/* clo_rng.in.h:52-62; `src` is the CUDA device header text of the generator */
struct clo_rng_info {
	const char* name;
	const char* src;
	const size_t seed_size;
};

/* clo_rng.in.h:67 / clo_rng.c:60-68 (NULL-terminated) */
extern const struct clo_rng_info clo_rng_infos[];

/* clo_rng.in.h:78-92 */
typedef enum clo_rng_seed_type {
	CLO_RNG_SEED_DEV_GID = 0,
	CLO_RNG_SEED_HOST_MT = 1,
	CLO_RNG_SEED_EXT_DEV = 2,
	CLO_RNG_SEED_EXT_HOST = 3
} CloRngSeedType;

/* clo_rng.in.h:97-99 / clo_rng.c:262-405.
 * `hash` (DEV_GID only) is an OpenCL C macro body in the reference
 * (clo_rng.c:101-109); precompiled here: NULL, "", "x" (no hash), "KNUTH(x)",
 * "XS1(x)" (clo_rng_init.cl:29-35); an expression without an assignment is a
 * no-op there too; any other string is built into the seeding kernel at run
 * time with NVRTC (one that does not compile fails with CLO_ERROR_ARGS). */
CloRng* clo_rng_new(const char* type, CloRngSeedType seed_type,
	void* seeds, size_t seeds_count, cl_ulong main_seed,
	const char* hash, CCLContext* ctx, CCLQueue* cq, GError** err);
/* clo_rng.c:412-426 */
void clo_rng_destroy(CloRng* rng);
/* clo_rng.c:438-446: CUDA device source (__device__ clo_rng_next + next_int API) */
const char* clo_rng_get_source(CloRng* rng);
/* clo_rng.c:456-463 */
CCLBuffer* clo_rng_get_device_seeds(CloRng* rng);
/* clo_rng.c:473-481 */
size_t clo_rng_get_size(CloRng* rng);

#ifdef __cplusplus
}
#endif
#endif
