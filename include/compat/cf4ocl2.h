/*
 * include/compat/cf4ocl2.h -- the cf4ocl2 handle types the cl_ops API is written
 * against, mapped onto CUDA:
 *
 *   CCLContext  -> one CUDA device ordinal (primary context)
 *   CCLDevice   -> device ordinal
 *   CCLQueue    -> cudaStream_t (+ a per-operation cudaEvent log when created
 *                  with CL_QUEUE_PROFILING_ENABLE)
 *   CCLBuffer   -> { device pointer, size, refcount, owner flag }
 *   CCLEvent    -> { cudaEvent_t start, end; name }  (owned by its queue)
 *   CCLProf     -> sum over a queue's event log
 *   CCLProgram  -> opaque tag (there is no run-time kernel compilation)
 *
 * cf4ocl2 is an un-vendored dependency of the reference (cmake/Modules/Findcf4ocl2.cmake:11-14);
 * only the symbols that cross the cl_ops API or that its callers need to create /
 * fill / read buffers are provided (call-site census: SURVEY.md section 8b).
 */
#ifndef CLO_B200_COMPAT_CF4OCL2_H
#define CLO_B200_COMPAT_CF4OCL2_H

#include <stddef.h>
#include <stdint.h>
#include "glib.h"

#ifdef __cplusplus
extern "C" {
#endif

/* OpenCL scalar typedefs and constants used by the API and its callers. */
typedef int8_t cl_char;
typedef uint8_t cl_uchar;
typedef int16_t cl_short;
typedef uint16_t cl_ushort;
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef int64_t cl_long;
typedef uint64_t cl_ulong;
typedef uint16_t cl_half;
typedef float cl_float;
typedef double cl_double;
typedef cl_uint cl_bool;
/* limits of the OpenCL scalar types (CL/cl_platform.h), used by src/benchmarks/clo_bench.c:33-142 */
#define CL_CHAR_MAX 127
#define CL_CHAR_MIN (-127 - 1)
#define CL_UCHAR_MAX 255
#define CL_SHRT_MAX 32767
#define CL_SHRT_MIN (-32767 - 1)
#define CL_USHRT_MAX 65535
#define CL_INT_MAX 2147483647
#define CL_INT_MIN (-2147483647 - 1)
#define CL_UINT_MAX 0xffffffffU
#define CL_LONG_MAX ((cl_long) 0x7FFFFFFFFFFFFFFFLL)
#define CL_LONG_MIN ((cl_long) -0x7FFFFFFFFFFFFFFFLL - 1LL)
#define CL_ULONG_MAX ((cl_ulong) 0xFFFFFFFFFFFFFFFFULL)
#define CL_FLT_MAX 340282346638528859811704183484516925440.0f
#define CL_FLT_MIN 1.175494350822287507969e-38f
#define CL_DBL_MAX 1.7976931348623157e308
#define CL_DBL_MIN 2.225073858507201383090e-308
typedef cl_ulong cl_mem_flags;
typedef cl_ulong cl_command_queue_properties;

#define CL_FALSE 0
#define CL_TRUE 1
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_QUEUE_PROFILING_ENABLE (1 << 1)

typedef struct ccl_context CCLContext;
typedef struct ccl_device CCLDevice;
typedef struct ccl_queue CCLQueue;
typedef struct ccl_buffer CCLBuffer;
typedef struct ccl_buffer CCLMemObj;
typedef struct ccl_event CCLEvent;
typedef struct ccl_program CCLProgram;
typedef struct ccl_prof CCLProf;
typedef CCLEvent** CCLEventWaitList;

/* ---- contexts / devices ---- */
CCLContext* ccl_context_new_any(GError** err);
CCLContext* ccl_context_new_gpu(GError** err);
/* dev_idx: pointer to an int device index, or NULL / -1 for "$CLO_DEVICE or 0"
 * (replaces the interactive menu of ccl_context_new_from_menu_full). */
CCLContext* ccl_context_new_from_menu_full(void* dev_idx, GError** err);
void ccl_context_ref(CCLContext* ctx);
void ccl_context_unref(CCLContext* ctx);
void ccl_context_destroy(CCLContext* ctx);
CCLDevice* ccl_context_get_device(CCLContext* ctx, cl_uint index, GError** err);
cl_uint ccl_context_get_num_devices(CCLContext* ctx, GError** err);
int ccl_device_get_ordinal(CCLDevice* dev);

/* ---- queues ---- */
CCLQueue* ccl_queue_new(CCLContext* ctx, CCLDevice* dev,
	cl_command_queue_properties properties, GError** err);
void ccl_queue_destroy(CCLQueue* cq);
cl_bool ccl_queue_finish(CCLQueue* cq, GError** err);
CCLDevice* ccl_queue_get_device(CCLQueue* cq, GError** err);
CCLContext* ccl_queue_get_context(CCLQueue* cq, GError** err);
/* drop the queue's event log (cf4ocl: ccl_queue_gc) */
void ccl_queue_gc(CCLQueue* cq);

/* ---- buffers ---- */
CCLBuffer* ccl_buffer_new(CCLContext* ctx, cl_mem_flags flags, size_t size,
	void* host_ptr, GError** err);
void ccl_buffer_ref(CCLBuffer* buf);
void ccl_buffer_destroy(CCLBuffer* buf);
CCLEvent* ccl_buffer_enqueue_write(CCLBuffer* buf, CCLQueue* cq,
	cl_bool blocking_write, size_t offset, size_t size, void* ptr,
	CCLEventWaitList* evt_wait_lst, GError** err);
CCLEvent* ccl_buffer_enqueue_read(CCLBuffer* buf, CCLQueue* cq,
	cl_bool blocking_read, size_t offset, size_t size, void* ptr,
	CCLEventWaitList* evt_wait_lst, GError** err);
CCLEvent* ccl_buffer_enqueue_copy(CCLBuffer* src_buf, CCLBuffer* dst_buf,
	CCLQueue* cq, size_t src_offset, size_t dst_offset, size_t size,
	CCLEventWaitList* evt_wait_lst, GError** err);
size_t ccl_memobj_get_size(CCLMemObj* mo, GError** err);

/* ---- events ---- */
void ccl_event_set_name(CCLEvent* evt, const char* name);
const char* ccl_event_get_name(CCLEvent* evt);
CCLEventWaitList* ccl_ewl(CCLEventWaitList* ewl, ...);
void ccl_event_wait_list_add(CCLEventWaitList* ewl, ...);
void ccl_event_wait_list_clear(CCLEventWaitList* ewl);
cl_bool ccl_event_wait(CCLEventWaitList* ewl, GError** err);

/* ---- profiler ---- */
CCLProf* ccl_prof_new(void);
void ccl_prof_destroy(CCLProf* prof);
void ccl_prof_add_queue(CCLProf* prof, const char* cq_name, CCLQueue* cq);
cl_bool ccl_prof_calc(CCLProf* prof, GError** err);
cl_ulong ccl_prof_get_duration(CCLProf* prof);

/* ---- leak check: true iff no shim handle is alive ---- */
cl_bool ccl_wrapper_memcheck(void);


/* programs and kernels built from source at run time (NVRTC): what clo_rng_bench.c:176-312 and
 * tests/test_rng.c:85-118 do with clo_rng_get_source() */
typedef struct ccl_kernel CCLKernel;
typedef struct ccl_arg CCLArg;
CCLProgram* ccl_program_new_from_source(CCLContext* ctx, const char* src, GError** err);
cl_bool ccl_program_build(CCLProgram* prg, const char* options, GError** err);
CCLKernel* ccl_program_get_kernel(CCLProgram* prg, const char* kernel_name, GError** err);
void ccl_program_destroy(CCLProgram* prg);
CCLArg* ccl_arg_new(void* value, size_t size);
#define ccl_arg_priv(value, type) ccl_arg_new(&value, sizeof(type))
void ccl_kernel_set_args(CCLKernel* krnl, ...);
CCLEvent* ccl_kernel_enqueue_ndrange(CCLKernel* krnl, CCLQueue* cq, cl_uint work_dim,
	const size_t* gwo, const size_t* gws, const size_t* lws, CCLEventWaitList* ewl, GError** err);
CCLEvent* ccl_kernel_set_args_and_enqueue_ndrange(CCLKernel* krnl, CCLQueue* cq, cl_uint work_dim,
	const size_t* gwo, const size_t* gws, const size_t* lws, CCLEventWaitList* ewl, GError** err, ...);
cl_bool ccl_kernel_suggest_worksizes(CCLKernel* krnl, CCLDevice* dev, cl_uint dims,
	const size_t* real_ws, size_t* gws, size_t* lws, GError** err);

/* ---- CUDA-side constructors / accessors (additive; not in cf4ocl2) ---- */
/* Wrap an existing cudaStream_t (passed as void*; NULL = legacy default stream). */
CCLQueue* ccl_queue_new_wrap(CCLContext* ctx, void* cuda_stream, GError** err);
void* ccl_queue_get_stream(CCLQueue* cq);
/* Wrap an existing device allocation; the buffer does not own the memory. */
CCLBuffer* ccl_buffer_new_wrap(CCLContext* ctx, void* device_ptr, size_t size, GError** err);
void* ccl_buffer_get_ptr(CCLBuffer* buf);
/* elapsed device time of one event in nanoseconds (needs a profiling queue) */
cl_ulong ccl_event_get_duration_ns(CCLEvent* evt, GError** err);

#ifdef __cplusplus
}
#endif
#endif
