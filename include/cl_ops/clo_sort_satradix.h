/*
 * clo_sort_satradix.h -- per-algorithm public header of the "satradix" sorter, as listed by the
 * reference's umbrella header (/root/reference/src/cl_ops/cl_ops.h:44-47; the reference's own is
 * src/cl_ops/sort/clo_sort_satradix.in.h:30-55).
 *
 * The reference runs {satradix_localsort, satradix_histogram, satradix_scatter} once per digit.
 * Here the tile sort and the scatter of a digit are ONE kernel (the onesweep pass) and the digit
 * histograms of all passes come from one read of the keys, so the kernel table reads:
 *   0  clo_radix_histogram     every digit's 256-bin histogram, one pass over the keys
 *   1  clo_radix_scan_bins     exclusive scan of the bins (global digit offsets)
 *   2  clo_radix_onesweep_v6   rank + decoupled prefix + staged scatter, once per 8-bit digit
 * The names are the CUDA kernels' real names (clo_sort_get_kernel_name returns them) and
 * clo_sort_get_localmem_usage returns their real shared-memory sizes.
 */
#ifndef CLO_B200_SORT_SATRADIX_H
#define CLO_B200_SORT_SATRADIX_H

#include <cl_ops/clo_sort_abstract.h>
#include <cl_ops/clo_scan_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLO_SORT_SATRADIX_NUM_KERNELS 3

#define CLO_SORT_SATRADIX_KIDX_HISTOGRAM 0
#define CLO_SORT_SATRADIX_KIDX_SCANBINS 1
#define CLO_SORT_SATRADIX_KIDX_ONESWEEP 2
/* the reference's indices for its tile sort and its scatter: both live in the onesweep pass */
#define CLO_SORT_SATRADIX_KIDX_LOCALSORT CLO_SORT_SATRADIX_KIDX_ONESWEEP
#define CLO_SORT_SATRADIX_KIDX_SCATTER CLO_SORT_SATRADIX_KIDX_ONESWEEP

#define CLO_SORT_SATRADIX_KNAME_HISTOGRAM "clo_radix_histogram"
#define CLO_SORT_SATRADIX_KNAME_SCANBINS "clo_radix_scan_bins"
#define CLO_SORT_SATRADIX_KNAME_ONESWEEP "clo_radix_onesweep_v6"
#define CLO_SORT_SATRADIX_KNAME_LOCALSORT CLO_SORT_SATRADIX_KNAME_ONESWEEP
#define CLO_SORT_SATRADIX_KNAME_SCATTER CLO_SORT_SATRADIX_KNAME_ONESWEEP

#define CLO_SORT_SATRADIX_KERNELNAMES { \
	CLO_SORT_SATRADIX_KNAME_HISTOGRAM, \
	CLO_SORT_SATRADIX_KNAME_SCANBINS, \
	CLO_SORT_SATRADIX_KNAME_ONESWEEP }

/* clo_sort_satradix.in.h:55 */
extern const CloSortImplDef clo_sort_satradix_def;

#ifdef __cplusplus
}
#endif
#endif
