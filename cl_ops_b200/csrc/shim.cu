/*
 * shim.cu -- cf4ocl2 / GLib stand-ins over the CUDA runtime (handles only).
 * Contract: include/compat/cf4ocl2.h, include/compat/glib.h.
 */
#include "clo_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_set>

std::atomic<unsigned long long> clo_launches{0};

/* ------------------------------------------------------------------ GLib */

static std::mutex g_quark_mtx;
static std::vector<std::string>& quark_table() {
	static std::vector<std::string> t{std::string("")};
	return t;
}

extern "C" GQuark g_quark_from_static_string(const gchar* string) {
	if (!string) return 0;
	std::lock_guard<std::mutex> lk(g_quark_mtx);
	auto& t = quark_table();
	for (size_t i = 1; i < t.size(); ++i)
		if (t[i] == string) return (GQuark) i;
	t.emplace_back(string);
	return (GQuark) (t.size() - 1);
}

extern "C" const gchar* g_quark_to_string(GQuark quark) {
	std::lock_guard<std::mutex> lk(g_quark_mtx);
	auto& t = quark_table();
	return (quark && quark < t.size()) ? t[quark].c_str() : NULL;
}

extern "C" void g_error_free(GError* error) {
	if (!error) return;
	free(error->message);
	free(error);
}

extern "C" void clo_b200_error_free(GError* error) { g_error_free(error); }

extern "C" void g_clear_error(GError** err) {
	if (err && *err) { g_error_free(*err); *err = NULL; }
}

extern "C" void g_set_error(GError** err, GQuark domain, gint code,
		const gchar* format, ...) {
	if (!err) return;
	if (*err) return; /* GLib warns and keeps the first error */
	char buf[1024];
	va_list ap;
	va_start(ap, format);
	vsnprintf(buf, sizeof(buf), format, ap);
	va_end(ap);
	GError* e = (GError*) malloc(sizeof(GError));
	e->domain = domain;
	e->code = code;
	e->message = strdup(buf);
	*err = e;
}

extern "C" void g_propagate_error(GError** dest, GError* src) {
	if (!src) return;
	if (!dest || *dest) { g_error_free(src); return; }
	*dest = src;
}

bool clo_cuda_failed(cudaError_t e, GError** err, const char* what) {
	if (e == cudaSuccess) return false;
	g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "CUDA error %d (%s) in %s",
		(int) e, cudaGetErrorString(e), what);
	cudaGetLastError(); /* clear sticky-less error state */
	return true;
}

/* -------------------------------------------------------------- registry */

static std::mutex g_handle_mtx;
static std::unordered_set<void*>& handle_set() {
	static std::unordered_set<void*> s;
	return s;
}

void clo_handle_add(void* h) {
	std::lock_guard<std::mutex> lk(g_handle_mtx);
	handle_set().insert(h);
}

bool clo_handle_alive(void* h) {
	std::lock_guard<std::mutex> lk(g_handle_mtx);
	return handle_set().count(h) != 0;
}

bool clo_handle_remove(void* h) {
	std::lock_guard<std::mutex> lk(g_handle_mtx);
	return handle_set().erase(h) != 0;
}

extern "C" cl_bool ccl_wrapper_memcheck(void) {
	std::lock_guard<std::mutex> lk(g_handle_mtx);
	return handle_set().empty() ? CL_TRUE : CL_FALSE;
}

int clo_sm_count(int ordinal) {
	static int cache[64];
	if (ordinal < 0 || ordinal >= 64) return 148;
	if (!cache[ordinal]) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, ordinal) != cudaSuccess || n <= 0)
			n = 148;
		cache[ordinal] = n;
	}
	return cache[ordinal];
}

/* --------------------------------------------------------------- context */

static CCLContext* context_for(int ordinal, GError** err) {
	int count = 0;
	if (clo_cuda_failed(cudaGetDeviceCount(&count), err, "cudaGetDeviceCount")) return NULL;
	if (count <= 0 || ordinal < 0 || ordinal >= count) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY,
			"No CUDA device with index %d (%d devices visible)", ordinal, count);
		return NULL;
	}
	ccl_context* ctx = new ccl_context();
	ctx->refs = 1;
	ctx->dev.ordinal = ordinal;
	clo_handle_add(ctx);
	return ctx;
}

static int default_ordinal() {
	const char* e = getenv("CLO_DEVICE");
	if (e && *e) return atoi(e);
	int cur = 0;
	if (cudaGetDevice(&cur) == cudaSuccess) return cur;
	return 0;
}

extern "C" CCLContext* ccl_context_new_any(GError** err) { return context_for(default_ordinal(), err); }
extern "C" CCLContext* ccl_context_new_gpu(GError** err) { return context_for(default_ordinal(), err); }

extern "C" CCLContext* ccl_context_new_from_menu_full(void* dev_idx, GError** err) {
	int idx = dev_idx ? *(int*) dev_idx : -1;
	if (idx < 0) idx = default_ordinal();
	return context_for(idx, err);
}

extern "C" void ccl_context_ref(CCLContext* ctx) { if (ctx && clo_handle_alive(ctx)) ctx->refs++; }

extern "C" void ccl_context_unref(CCLContext* ctx) {
	if (!ctx || !clo_handle_alive(ctx)) return;
	if (--ctx->refs <= 0) { clo_handle_remove(ctx); delete ctx; }
}

extern "C" void ccl_context_destroy(CCLContext* ctx) { ccl_context_unref(ctx); }

extern "C" CCLDevice* ccl_context_get_device(CCLContext* ctx, cl_uint index, GError** err) {
	if (!ctx || index != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Device index %u out of range", index);
		return NULL;
	}
	return &ctx->dev;
}

extern "C" cl_uint ccl_context_get_num_devices(CCLContext* ctx, GError** err) {
	(void) err;
	return ctx ? 1 : 0;
}

extern "C" int ccl_device_get_ordinal(CCLDevice* dev) { return dev ? dev->ordinal : -1; }

/* ----------------------------------------------------------------- queue */

extern "C" CCLQueue* ccl_queue_new(CCLContext* ctx, CCLDevice* dev,
		cl_command_queue_properties properties, GError** err) {
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	(void) dev;
	CloDeviceGuard g(ctx->dev.ordinal);
	cudaStream_t s;
	if (clo_cuda_failed(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), err, "cudaStreamCreate")) return NULL;
	ccl_queue* cq = new ccl_queue();
	cq->ctx = ctx; ccl_context_ref(ctx);
	cq->stream = s; cq->owns_stream = true;
	cq->profiling = (properties & CL_QUEUE_PROFILING_ENABLE) != 0;
	clo_handle_add(cq);
	return cq;
}

extern "C" CCLQueue* ccl_queue_new_wrap(CCLContext* ctx, void* cuda_stream, GError** err) {
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	ccl_queue* cq = new ccl_queue();
	cq->ctx = ctx; ccl_context_ref(ctx);
	cq->stream = (cudaStream_t) cuda_stream; cq->owns_stream = false;
	cq->profiling = false;
	clo_handle_add(cq);
	return cq;
}

extern "C" void* ccl_queue_get_stream(CCLQueue* cq) { return cq ? (void*) cq->stream : NULL; }

static void free_event(ccl_event* e) {
	if (e->start) cudaEventDestroy(e->start);
	if (e->end) cudaEventDestroy(e->end);
	delete e;
}

extern "C" void ccl_queue_gc(CCLQueue* cq) {
	if (!cq) return;
	for (ccl_event* e : cq->events) free_event(e);
	cq->events.clear();
}

extern "C" void ccl_queue_destroy(CCLQueue* cq) {
	if (!cq || !clo_handle_remove(cq)) return;
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	cudaStreamSynchronize(cq->stream);
	ccl_queue_gc(cq);
	if (cq->owns_stream) cudaStreamDestroy(cq->stream);
	ccl_context_unref(cq->ctx);
	delete cq;
}

extern "C" cl_bool ccl_queue_finish(CCLQueue* cq, GError** err) {
	if (!cq) return CL_FALSE;
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	return clo_cuda_failed(cudaStreamSynchronize(cq->stream), err, "cudaStreamSynchronize") ? CL_FALSE : CL_TRUE;
}

extern "C" CCLDevice* ccl_queue_get_device(CCLQueue* cq, GError** err) { (void) err; return cq ? &cq->ctx->dev : NULL; }
extern "C" CCLContext* ccl_queue_get_context(CCLQueue* cq, GError** err) { (void) err; return cq ? cq->ctx : NULL; }

/* The event log would grow without bound on a long-lived non-profiling queue;
 * keep only the most recent events there (profiling queues keep everything
 * until ccl_prof_calc / ccl_queue_gc, as cf4ocl does). */
static const size_t CLO_EVENT_LOG_CAP = 64;

ccl_event* clo_queue_begin(ccl_queue* cq, const char* name) {
	if (!cq->profiling && cq->events.size() >= CLO_EVENT_LOG_CAP && (cq->events.size() % CLO_EVENT_LOG_CAP) == 0) {
		/* Every CCLEvent stays a valid handle until ccl_queue_gc / ccl_queue_destroy, as in cf4ocl
		 * (a client may keep old events in wait lists).  What is given back early is only the CUDA
		 * event of entries that have already completed: an entry without one reads as signalled
		 * (ccl_event_wait and the wait lists skip it). */
		for (ccl_event* old : cq->events)
			if (old->end && cudaEventQuery(old->end) == cudaSuccess) {
				cudaEventDestroy(old->end);
				old->end = nullptr;
			}
	}
	ccl_event* e = new ccl_event();
	e->start = nullptr; e->end = nullptr; e->cq = cq;
	e->name = name ? name : "";
	if (cq->profiling) {
		cudaEventCreate(&e->start);
		cudaEventRecord(e->start, cq->stream);
	}
	return e;
}

void clo_queue_end(ccl_queue* cq, ccl_event* e) {
	if (cq->profiling) cudaEventCreate(&e->end);
	else cudaEventCreateWithFlags(&e->end, cudaEventDisableTiming);
	cudaEventRecord(e->end, cq->stream);
	cq->events.push_back(e);
}

/* ---------------------------------------------------------------- buffer */

extern "C" CCLBuffer* ccl_buffer_new(CCLContext* ctx, cl_mem_flags flags, size_t size,
		void* host_ptr, GError** err) {
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	CloDeviceGuard g(ctx->dev.ordinal);
	void* p = NULL;
	if (clo_cuda_failed(cudaMalloc(&p, size ? size : 1), err, "cudaMalloc")) return NULL;
	if ((flags & CL_MEM_COPY_HOST_PTR) && host_ptr && size) {
		if (clo_cuda_failed(cudaMemcpy(p, host_ptr, size, cudaMemcpyHostToDevice), err, "cudaMemcpy")) {
			cudaFree(p);
			return NULL;
		}
	}
	ccl_buffer* b = new ccl_buffer();
	b->refs = 1; b->ctx = ctx; b->ptr = p; b->size = size; b->owns = true;
	ccl_context_ref(ctx);
	clo_handle_add(b);
	return b;
}

extern "C" CCLBuffer* ccl_buffer_new_wrap(CCLContext* ctx, void* device_ptr, size_t size, GError** err) {
	if (!ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context"); return NULL; }
	ccl_buffer* b = new ccl_buffer();
	b->refs = 1; b->ctx = ctx; b->ptr = device_ptr; b->size = size; b->owns = false;
	ccl_context_ref(ctx);
	clo_handle_add(b);
	return b;
}

extern "C" void* ccl_buffer_get_ptr(CCLBuffer* buf) { return buf ? buf->ptr : NULL; }

extern "C" void ccl_buffer_ref(CCLBuffer* buf) { if (buf && clo_handle_alive(buf)) buf->refs++; }

extern "C" void ccl_buffer_destroy(CCLBuffer* buf) {
	/* double destroy must be harmless (clo_rng_bench.c:364,382) */
	if (!buf || !clo_handle_alive(buf)) return;
	if (--buf->refs > 0) return;
	clo_handle_remove(buf);
	if (buf->owns && buf->ptr) {
		CloDeviceGuard g(buf->ctx->dev.ordinal);
		cudaFree(buf->ptr);
	} else if (buf->ipc && buf->ptr) {
		CloDeviceGuard g(buf->ctx->dev.ordinal);
		cudaIpcCloseMemHandle(buf->ptr);
	}
	ccl_context_unref(buf->ctx);
	delete buf;
}

/* ---- peer memory of another process on the same box (one process per GPU): the receive
 *      buffers of the sample sort are exported once and written by the peers' scatter kernels */
extern "C" cl_bool clo_b200_ipc_export(CCLBuffer* buf, unsigned char handle[64], GError** err) {
	if (!buf || !buf->owns || !handle) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ipc export: need a buffer created by ccl_buffer_new"); return CL_FALSE; }
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	CloDeviceGuard g(buf->ctx->dev.ordinal);
	cudaIpcMemHandle_t h;
	if (clo_cuda_failed(cudaIpcGetMemHandle(&h, buf->ptr), err, "cudaIpcGetMemHandle")) return CL_FALSE;
	memcpy(handle, &h, 64);
	return CL_TRUE;
}

extern "C" CCLBuffer* clo_b200_ipc_import(CCLContext* ctx, const unsigned char handle[64], size_t size, GError** err) {
	if (!ctx || !handle) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ipc import: NULL argument"); return NULL; }
	CloDeviceGuard g(ctx->dev.ordinal);
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	void* p = NULL;
	if (clo_cuda_failed(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), err, "cudaIpcOpenMemHandle")) return NULL;
	ccl_buffer* b = new ccl_buffer();
	b->refs = 1; b->ctx = ctx; b->ptr = p; b->size = size; b->owns = false; b->ipc = true;
	ccl_context_ref(ctx);
	clo_handle_add(b);
	return b;
}

extern "C" size_t ccl_memobj_get_size(CCLMemObj* mo, GError** err) { (void) err; return mo ? mo->size : 0; }

static void consume_ewl(CCLQueue* cq, CCLEventWaitList* ewl) {
	if (!ewl || !*ewl) return;
	for (CCLEvent** p = *ewl; *p; ++p)
		if ((*p)->end && (*p)->cq != cq) cudaStreamWaitEvent(cq->stream, (*p)->end, 0);
	ccl_event_wait_list_clear(ewl);
}

static CCLEvent* enqueue_copy(CCLQueue* cq, void* dst, const void* src, size_t size,
		cudaMemcpyKind kind, bool blocking, const char* name, CCLEventWaitList* ewl, GError** err) {
	if (!cq) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL queue"); return NULL; }
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	consume_ewl(cq, ewl);
	ccl_event* e = clo_queue_begin(cq, name);
	cudaError_t rc = size ? cudaMemcpyAsync(dst, src, size, kind, cq->stream) : cudaSuccess;
	clo_queue_end(cq, e);
	if (clo_cuda_failed(rc, err, name)) return NULL;
	if (blocking && clo_cuda_failed(cudaStreamSynchronize(cq->stream), err, name)) return NULL;
	return e;
}

extern "C" CCLEvent* ccl_buffer_enqueue_write(CCLBuffer* buf, CCLQueue* cq, cl_bool blocking_write,
		size_t offset, size_t size, void* ptr, CCLEventWaitList* ewl, GError** err) {
	if (!buf || offset + size > buf->size) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Write of %zu bytes at %zu exceeds buffer", size, offset);
		return NULL;
	}
	return enqueue_copy(cq, (char*) buf->ptr + offset, ptr, size, cudaMemcpyHostToDevice,
		blocking_write != 0, "write_buffer", ewl, err);
}

extern "C" CCLEvent* ccl_buffer_enqueue_read(CCLBuffer* buf, CCLQueue* cq, cl_bool blocking_read,
		size_t offset, size_t size, void* ptr, CCLEventWaitList* ewl, GError** err) {
	if (!buf || offset + size > buf->size) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Read of %zu bytes at %zu exceeds buffer", size, offset);
		return NULL;
	}
	return enqueue_copy(cq, ptr, (char*) buf->ptr + offset, size, cudaMemcpyDeviceToHost,
		blocking_read != 0, "read_buffer", ewl, err);
}

extern "C" CCLEvent* ccl_buffer_enqueue_copy(CCLBuffer* src_buf, CCLBuffer* dst_buf, CCLQueue* cq,
		size_t src_offset, size_t dst_offset, size_t size, CCLEventWaitList* ewl, GError** err) {
	if (!src_buf || !dst_buf || src_offset + size > src_buf->size || dst_offset + size > dst_buf->size) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Copy of %zu bytes exceeds a buffer", size);
		return NULL;
	}
	return enqueue_copy(cq, (char*) dst_buf->ptr + dst_offset, (char*) src_buf->ptr + src_offset, size,
		cudaMemcpyDeviceToDevice, false, "copy_buffer", ewl, err);
}

/* ---------------------------------------------------------------- events */

extern "C" void ccl_event_set_name(CCLEvent* evt, const char* name) { if (evt && name) evt->name = name; }
extern "C" const char* ccl_event_get_name(CCLEvent* evt) { return evt ? evt->name.c_str() : NULL; }

extern "C" void ccl_event_wait_list_clear(CCLEventWaitList* ewl) {
	if (ewl && *ewl) { free(*ewl); *ewl = NULL; }
}

static void ewl_add_v(CCLEventWaitList* ewl, va_list ap) {
	size_t n = 0;
	if (*ewl) while ((*ewl)[n]) ++n;
	for (;;) {
		CCLEvent* e = va_arg(ap, CCLEvent*);
		if (!e) break;
		*ewl = (CCLEvent**) realloc(*ewl, (n + 2) * sizeof(CCLEvent*));
		(*ewl)[n++] = e;
		(*ewl)[n] = NULL;
	}
}

extern "C" CCLEventWaitList* ccl_ewl(CCLEventWaitList* ewl, ...) {
	va_list ap;
	va_start(ap, ewl);
	ewl_add_v(ewl, ap);
	va_end(ap);
	return ewl;
}

extern "C" void ccl_event_wait_list_add(CCLEventWaitList* ewl, ...) {
	va_list ap;
	va_start(ap, ewl);
	ewl_add_v(ewl, ap);
	va_end(ap);
}

extern "C" cl_bool ccl_event_wait(CCLEventWaitList* ewl, GError** err) {
	cl_bool ok = CL_TRUE;
	if (ewl && *ewl) {
		for (CCLEvent** p = *ewl; *p; ++p)
			if ((*p)->end && clo_cuda_failed(cudaEventSynchronize((*p)->end), err, "cudaEventSynchronize")) ok = CL_FALSE;
		ccl_event_wait_list_clear(ewl);
	}
	return ok;
}

extern "C" cl_ulong ccl_event_get_duration_ns(CCLEvent* evt, GError** err) {
	if (!evt || !evt->start || !evt->end) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Event has no timing (queue created without CL_QUEUE_PROFILING_ENABLE)");
		return 0;
	}
	float ms = 0;
	if (clo_cuda_failed(cudaEventSynchronize(evt->end), err, "cudaEventSynchronize")) return 0;
	if (clo_cuda_failed(cudaEventElapsedTime(&ms, evt->start, evt->end), err, "cudaEventElapsedTime")) return 0;
	return (cl_ulong) ((double) ms * 1e6);
}

/* -------------------------------------------------------------- profiler */

extern "C" CCLProf* ccl_prof_new(void) {
	ccl_prof* p = new ccl_prof();
	p->duration_ns = 0;
	clo_handle_add(p);
	return p;
}

extern "C" void ccl_prof_destroy(CCLProf* prof) {
	if (!prof || !clo_handle_remove(prof)) return;
	delete prof;
}

extern "C" void ccl_prof_add_queue(CCLProf* prof, const char* cq_name, CCLQueue* cq) {
	(void) cq_name;
	if (prof && cq) prof->queues.push_back(cq);
}

/* Sum of the device time of every logged operation; drains the logs
 * (cf4ocl's ccl_prof_calc consumes the queue's events the same way). */
extern "C" cl_bool ccl_prof_calc(CCLProf* prof, GError** err) {
	if (!prof) return CL_FALSE;
	double total_ms = 0;
	for (ccl_queue* cq : prof->queues) {
		CloDeviceGuard g(cq->ctx->dev.ordinal);
		if (clo_cuda_failed(cudaStreamSynchronize(cq->stream), err, "cudaStreamSynchronize")) return CL_FALSE;
		for (ccl_event* e : cq->events) {
			if (e->start && e->end) {
				float ms = 0;
				if (cudaEventElapsedTime(&ms, e->start, e->end) == cudaSuccess) total_ms += ms;
			}
		}
		ccl_queue_gc(cq);
	}
	prof->duration_ns = (unsigned long long) (total_ms * 1e6);
	return CL_TRUE;
}

extern "C" cl_ulong ccl_prof_get_duration(CCLProf* prof) { return prof ? prof->duration_ns : 0; }
