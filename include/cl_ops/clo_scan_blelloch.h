/*
 * clo_scan_blelloch.h -- per-algorithm public header of the "blelloch" scanner
 * (/root/reference/src/cl_ops/cl_ops.h:50-52; reference: src/cl_ops/scan/clo_scan_blelloch.in.h:30-46).
 *
 * The reference runs three kernels (workgroupScan, workgroupSumsScan, addWorkgroupSums) and moves
 * 16-24 bytes per element.  Here the exclusive scan is a single pass -- every element read once
 * and written once:
 *   0  clo_scan_tma       element and sum type of the same size: persistent CTAs, tiles moved by the
 *                         copy engine (cp.async.bulk.tensor + mbarrier ring, SWIZZLE_128B), scanned in
 *                         place in shared memory, prefix-propagator CTA (a chain of warps)
 *   1  clo_scan_pp        other type pairs: persistent CTAs, cp.async tile ring, same propagator
 *   2  clo_scan_lookback  one tile per CTA, decoupled look-back (small or unaligned inputs)
 *   3  clo_scan_reduce_partial / 4 clo_scan_reduce_final   the per-GPU total of the multi-GPU
 *      scan (carry-in of the GPUs behind)
 */
#ifndef CLO_B200_SCAN_BLELLOCH_H
#define CLO_B200_SCAN_BLELLOCH_H

#include <cl_ops/clo_scan_abstract.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLO_SCAN_BLELLOCH_NUM_KERNELS 5
#define CLO_SCAN_BLELLOCH_KIDX_TMA 0
#define CLO_SCAN_BLELLOCH_KIDX_PP 1
#define CLO_SCAN_BLELLOCH_KIDX_LOOKBACK 2
#define CLO_SCAN_BLELLOCH_KIDX_REDUCE_PARTIAL 3
#define CLO_SCAN_BLELLOCH_KIDX_REDUCE_FINAL 4
/* the reference's per-work-group scan is where the elements are read and written: the
 * single-pass kernel; its two fix-up kernels have no counterpart */
#define CLO_SCAN_BLELLOCH_KIDX_WGSCAN CLO_SCAN_BLELLOCH_KIDX_TMA
#define CLO_SCAN_BLELLOCH_KNAME_TMA "clo_scan_tma"
#define CLO_SCAN_BLELLOCH_KNAME_PP "clo_scan_pp"
#define CLO_SCAN_BLELLOCH_KNAME_LOOKBACK "clo_scan_lookback"
#define CLO_SCAN_BLELLOCH_KNAME_REDUCE_PARTIAL "clo_scan_reduce_partial"
#define CLO_SCAN_BLELLOCH_KNAME_REDUCE_FINAL "clo_scan_reduce_final"
#define CLO_SCAN_BLELLOCH_KNAME_WGSCAN CLO_SCAN_BLELLOCH_KNAME_TMA

/* clo_scan_blelloch.in.h:46 */
extern const CloScanImplDef clo_scan_blelloch_def;

#ifdef __cplusplus
}
#endif
#endif
