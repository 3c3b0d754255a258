/*
 * include/cl_ops/clo_scan_abstract.h -- the scanner object.
 * Replaces: /root/reference/src/cl_ops/scan/clo_scan_abstract.in.h:41-162
 *
 * Algorithm name: "blelloch" (clo_scan_abstract.c:86-89).  Semantics: exclusive
 * prefix sum, out[0] = 0, out[i] = sum_{j<i} (SUM)in[j] in SUM-type arithmetic
 * (clo_scan_blelloch.cl:66-125).  Here: one single-pass decoupled-look-back kernel.
 */
#ifndef CLO_B200_SCAN_ABSTRACT_H
#define CLO_B200_SCAN_ABSTRACT_H

#include <cl_ops/clo_common.h>

#ifdef __cplusplus
extern "C" {
#endif

/* clo_scan_abstract.in.h:41-103 */
typedef struct clo_scan_impl_def {
	const char* name;
	const char* (*init)(CloScan* scanner, const char* options, GError** err);
	void (*finalize)(CloScan* scanner);
	CCLEvent* (*scan_with_device_data)(CloScan* scanner, CCLQueue* cq_exec,
		CCLQueue* cq_comm, CCLBuffer* data_in, CCLBuffer* data_out,
		size_t numel, size_t lws_max, GError** err);
	cl_uint (*get_num_kernels)(CloScan* scanner, GError** err);
	const char* (*get_kernel_name)(CloScan* scanner, cl_uint i, GError** err);
	size_t (*get_localmem_usage)(CloScan* scanner, cl_uint i, size_t lws_max,
		size_t numel, GError** err);
} CloScanImplDef;

/* clo_scan_abstract.in.h:109-111 / clo_scan_abstract.c:74-168 (types by value) */
CloScan* clo_scan_new(const char* type, const char* options,
	CCLContext* ctx, CloType elem_type, CloType sum_type,
	const char* compiler_opts, GError** err);
/* clo_scan_abstract.c:175-193 */
void clo_scan_destroy(CloScan* scan);
/* clo_scan_abstract.c:213-229 */
CCLEvent* clo_scan_with_device_data(CloScan* scanner, CCLQueue* cq_exec,
	CCLQueue* cq_comm, CCLBuffer* data_in, CCLBuffer* data_out,
	size_t numel, size_t lws_max, GError** err);
/* clo_scan_abstract.c:255-362 */
cl_bool clo_scan_with_host_data(CloScan* scanner, CCLQueue* cq_exec,
	CCLQueue* cq_comm, void* data_in, void* data_out, size_t numel,
	size_t lws_max, GError** err);
/* clo_scan_abstract.c:371-567 */
CCLContext* clo_scan_get_context(CloScan* scanner);
CCLProgram* clo_scan_get_program(CloScan* scanner);
CloType clo_scan_get_elem_type(CloScan* scanner);
size_t clo_scan_get_element_size(CloScan* scanner);
CloType clo_scan_get_sum_type(CloScan* scanner);
size_t clo_scan_get_sum_size(CloScan* scanner);
void* clo_scan_get_data(CloScan* scanner);
void clo_scan_set_data(CloScan* scanner, void* data);
cl_uint clo_scan_get_num_kernels(CloScan* scanner, GError** err);
const char* clo_scan_get_kernel_name(CloScan* scanner, cl_uint i, GError** err);
size_t clo_scan_get_localmem_usage(CloScan* scanner, cl_uint i,
	size_t lws_max, size_t numel, GError** err);

/* clo_scan_blelloch.in.h:46 */
extern const CloScanImplDef clo_scan_blelloch_def;

#ifdef __cplusplus
}
#endif
#endif
