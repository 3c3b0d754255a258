/*
 * include/cl_ops/clo_sort_abstract.h -- the sorter object.
 * Replaces: /root/reference/src/cl_ops/sort/clo_sort_abstract.in.h:43-170
 *
 * Algorithm names accepted by clo_sort_new (clo_sort_abstract.in.h:30):
 *   "satradix"  LSD radix sort on raw key bits, ascending, stable.  Here: 8-bit-digit
 *               onesweep (one histogram pass + one decoupled-look-back pass per digit).
 *   "sbitonic", "abitonic"  the canonical bitonic network (clo_sort_sbitonic.cl:38-69);
 *               identical results, honours `compare` and `get_key`.
 *   "gselect"   stable rank sort (clo_sort_gselect.cl:38-57), out of place.
 *
 * `compare` and `get_key` are OpenCL C macro bodies in the reference, spliced into its kernels
 * at build time (clo_sort_abstract.c:157-168).  Here a menu is precompiled -- compare in
 * { NULL, "((a) > (b))", "((a) < (b))" }; get_key in { NULL, "(x)", "((x) >> K)", "((x) & M)",
 * "(((x) >> K) & M)" } -- and ANY other string is compiled at run time with NVRTC, for all four
 * algorithms (satradix: the key extraction; its kernels never expand `compare`).  A string that
 * does not compile fails with CLO_ERROR_ARGS carrying the compiler log.
 */
#ifndef CLO_B200_SORT_ABSTRACT_H
#define CLO_B200_SORT_ABSTRACT_H

#include <cl_ops/clo_common.h>

#ifdef __cplusplus
extern "C" {
#endif

/* clo_sort_abstract.in.h:43-110 */
typedef struct clo_sort_impl_def {
	const char* name;
	cl_bool in_place;
	const char* (*init)(CloSort* sorter, const char* options, GError** err);
	void (*finalize)(CloSort* sorter);
	CCLEvent* (*sort_with_device_data)(CloSort* sorter, CCLQueue* cq_exec,
		CCLQueue* cq_comm, CCLBuffer* data_in, CCLBuffer* data_out,
		size_t numel, size_t lws_max, GError** err);
	cl_uint (*get_num_kernels)(CloSort* sorter, GError** err);
	const char* (*get_kernel_name)(CloSort* sorter, cl_uint i, GError** err);
	size_t (*get_localmem_usage)(CloSort* sorter, cl_uint i, size_t lws_max,
		size_t numel, GError** err);
} CloSortImplDef;

/* clo_sort_abstract.in.h:116-120 / clo_sort_abstract.c:91-207 */
CloSort* clo_sort_new(const char* type, const char* options,
	CCLContext* ctx, CloType* elem_type, CloType* key_type,
	const char* compare, const char* get_key, const char* compiler_opts,
	GError** err);
/* clo_sort_abstract.c:214-232 */
void clo_sort_destroy(CloSort* sorter);
/* clo_sort_abstract.c:252-270: data_out == NULL sorts data_in in place */
CCLEvent* clo_sort_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLQueue* cq_comm, CCLBuffer* data_in, CCLBuffer* data_out,
	size_t numel, size_t lws_max, GError** err);
/* clo_sort_abstract.c:296-418: alloc, H2D, sort, D2H; blocks until done */
cl_bool clo_sort_with_host_data(CloSort* sorter, CCLQueue* cq_exec,
	CCLQueue* cq_comm, void* data_in, void* data_out, size_t numel,
	size_t lws_max, GError** err);
/* clo_sort_abstract.c:427-629 */
CCLContext* clo_sort_get_context(CloSort* sorter);
CCLProgram* clo_sort_get_program(CloSort* sorter);
CloType clo_sort_get_element_type(CloSort* sorter);
size_t clo_sort_get_element_size(CloSort* sorter);
CloType clo_sort_get_key_type(CloSort* sorter);
size_t clo_sort_get_key_size(CloSort* sorter);
void* clo_sort_get_data(CloSort* sorter);
void clo_sort_set_data(CloSort* sorter, void* data);
cl_uint clo_sort_get_num_kernels(CloSort* sorter, GError** err);
const char* clo_sort_get_kernel_name(CloSort* sorter, cl_uint i, GError** err);
size_t clo_sort_get_localmem_usage(CloSort* sorter, cl_uint i,
	size_t lws_max, size_t numel, GError** err);

/* clo_sort_sbitonic.in.h:36, clo_sort_abitonic.in.h:116,
 * clo_sort_gselect.in.h:36, clo_sort_satradix.in.h:55 */
extern const CloSortImplDef clo_sort_sbitonic_def;
extern const CloSortImplDef clo_sort_abitonic_def;
extern const CloSortImplDef clo_sort_gselect_def;
extern const CloSortImplDef clo_sort_satradix_def;

#ifdef __cplusplus
}
#endif
#endif
