/*
 * clo_sort_sbitonic.h -- per-algorithm public header of the "sbitonic" sorter
 * (/root/reference/src/cl_ops/cl_ops.h:38-39; reference: src/cl_ops/sort/clo_sort_sbitonic.in.h:30-36).
 *
 * The reference launches its one "sbitonic" kernel once per step of the network (210 launches
 * for 2^20 keys).  Here the same network runs as one persistent launch (clo_bitonic_fused) when
 * N is a power of two, and as clo_bitonic_local / clo_bitonic_global on a padded copy otherwise.
 */
#ifndef CLO_B200_SORT_SBITONIC_H
#define CLO_B200_SORT_SBITONIC_H

#include <cl_ops/clo_sort_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLO_SORT_SBITONIC_NUM_KERNELS 3
#define CLO_SORT_SBITONIC_KIDX_FUSED 0
#define CLO_SORT_SBITONIC_KIDX_LOCAL 1
#define CLO_SORT_SBITONIC_KIDX_GLOBAL 2
#define CLO_SORT_SBITONIC_KNAME_FUSED "clo_bitonic_fused"
#define CLO_SORT_SBITONIC_KNAME_LOCAL "clo_bitonic_local"
#define CLO_SORT_SBITONIC_KNAME_GLOBAL "clo_bitonic_global"
/* the reference's single kernel name maps onto the kernel that runs the whole network */
#define CLO_SORT_SBITONIC_KNAME CLO_SORT_SBITONIC_KNAME_FUSED

/* clo_sort_sbitonic.in.h:36 */
extern const CloSortImplDef clo_sort_sbitonic_def;

#ifdef __cplusplus
}
#endif
#endif
