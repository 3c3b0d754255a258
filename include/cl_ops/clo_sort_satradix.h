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
 * clo_sort_get_localmem_usage returns their real shared-memory sizes.  Small helpers run beside
 * them: clo_radix_chain_fixup (one conditional copy after the last pass: a pass whose digit is
 * the same for every key moves nothing), clo_radix_typed_flip (typed_order), and for a get_key
 * string outside the menu the run-time compiled clo_jit_extract_keys + clo_radix_gather.
 *
 * Options (clo_sort_new's `options` string, comma separated key=value):
 *   radix=R        the reference's digit width option (clo_sort_satradix.c:352,385-392): a power
 *                  of two; only decides how many bits of the key are sorted, as there
 *                  (total_digits = elem_bits / log2 R); the passes here are always 8 bits wide.
 *   scan=blelloch, scan...   accepted as in the reference (it forwards them to its scanner).
 *   typed_order=1  NOT in the reference, opt-in: signed integers and floats sort by VALUE.  The
 *                  default is the reference's order: raw key bits ascending (negative integers
 *                  after positive ones), float keys rejected (`key >> b` does not compile there).
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
