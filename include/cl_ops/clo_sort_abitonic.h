/*
 * clo_sort_abitonic.h -- per-algorithm public header of the "abitonic" sorter
 * (/root/reference/src/cl_ops/cl_ops.h:35-36; reference: src/cl_ops/sort/clo_sort_abitonic.in.h:30-116).
 *
 * The reference plans a sequence out of 26 OpenCL kernels that fuse 1-4 steps of the bitonic
 * network in private memory (abit_priv_*), up to 11 in local memory (abit_local_s*), or both
 * (abit_hyb_*), steered by the options minps / maxps / maxsfs (clo_sort_abitonic.c:58-313,
 * 486-542).  Here ONE kernel template does all of that -- k <= maxps steps in registers per pass,
 * passes through a shared-memory tile of at most 2^min(maxsfs, 13) elements, grid barriers for
 * the larger strides -- so the kernel table is the sbitonic one and the options keep their
 * meaning as fusion depths:
 *   maxps   steps fused in registers per pass (1..4; the reference's 2s4v / 3s8v / 4s16v)
 *   minps   accepted and validated (minps <= maxps); a pass never fuses fewer steps than remain
 *   maxsfs  largest number of steps taken inside the shared-memory tile (the reference's s2..s11)
 */
#ifndef CLO_B200_SORT_ABITONIC_H
#define CLO_B200_SORT_ABITONIC_H

#include <cl_ops/clo_sort_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLO_SORT_ABITONIC_NUM_KERNELS 3
#define CLO_SORT_ABITONIC_KIDX_FUSED 0
#define CLO_SORT_ABITONIC_KIDX_LOCAL 1
#define CLO_SORT_ABITONIC_KIDX_GLOBAL 2
/* the reference's "any step" kernel index: the fused kernel takes every step */
#define CLO_SORT_ABITONIC_KIDX_ANY CLO_SORT_ABITONIC_KIDX_FUSED
#define CLO_SORT_ABITONIC_KNAME_FUSED "clo_bitonic_fused"
#define CLO_SORT_ABITONIC_KNAME_LOCAL "clo_bitonic_local"
#define CLO_SORT_ABITONIC_KNAME_GLOBAL "clo_bitonic_global"
#define CLO_SORT_ABITONIC_KNAME_ANY CLO_SORT_ABITONIC_KNAME_FUSED

/* clo_sort_abitonic.in.h:116 */
extern const CloSortImplDef clo_sort_abitonic_def;

#ifdef __cplusplus
}
#endif
#endif
