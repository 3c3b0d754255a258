/*
 * include/cl_ops.h -- aggregate header of the cl_ops API (B200 backend).
 * Replaces: /root/reference/src/cl_ops/cl_ops.h:29-52
 */
#ifndef CLO_B200_CL_OPS_H
#define CLO_B200_CL_OPS_H

#ifdef __cplusplus
extern "C" {
#endif

#include <cl_ops/clo_common.h>
#include <cl_ops/clo_rng.h>
#include <cl_ops/clo_sort_abstract.h>
#include <cl_ops/clo_sort_abitonic.h>
#include <cl_ops/clo_sort_sbitonic.h>
#include <cl_ops/clo_sort_gselect.h>
#include <cl_ops/clo_sort_satradix.h>
#include <cl_ops/clo_scan_abstract.h>
#include <cl_ops/clo_scan_blelloch.h>
#include <cl_ops/clo_b200.h>

#ifdef __cplusplus
}
#endif
#endif
