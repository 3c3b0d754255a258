/*
 * clo_sort_gselect.h -- per-algorithm public header of the "gselect" sorter
 * (/root/reference/src/cl_ops/cl_ops.h:41-42; reference: src/cl_ops/sort/clo_sort_gselect.in.h:30-36).
 * One kernel, the stable rank sort of clo_sort_gselect.cl:38-57 with a shared-memory key tile.
 */
#ifndef CLO_B200_SORT_GSELECT_H
#define CLO_B200_SORT_GSELECT_H

#include <cl_ops/clo_sort_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLO_SORT_GSELECT_NUM_KERNELS 1
#define CLO_SORT_GSELECT_KNAME "clo_gselect_kernel"

/* clo_sort_gselect.in.h:36 */
extern const CloSortImplDef clo_sort_gselect_def;

#ifdef __cplusplus
}
#endif
#endif
