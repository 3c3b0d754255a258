/*
 * clo_internal.h -- private definitions shared by the host side of the library:
 * the concrete structs behind the cf4ocl2-style handles, CUDA error plumbing and
 * the launch counter.  Not installed.
 */
#ifndef CLO_INTERNAL_H
#define CLO_INTERNAL_H

#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include <cl_ops.h>

/* ---- handle structs (see include/compat/cf4ocl2.h for the mapping) ---- */

struct ccl_device {
	int ordinal;
};

struct ccl_context {
	int refs;
	ccl_device dev;
};

struct ccl_event {
	cudaEvent_t start;   /* nullptr unless the queue profiles */
	cudaEvent_t end;
	std::string name;
	struct ccl_queue* cq;
};

struct ccl_queue {
	ccl_context* ctx;
	cudaStream_t stream;
	bool owns_stream;
	bool profiling;
	std::vector<ccl_event*> events;   /* the queue owns its events */
};

struct ccl_buffer {
	int refs;
	ccl_context* ctx;
	void* ptr;
	size_t size;
	bool owns;
	bool ipc = false;      /* ptr came from cudaIpcOpenMemHandle */
};

struct ccl_kernel;
struct ccl_program {
	const char* tag;
	/* run-time compiled programs (ccl_program_new_from_source, jit.cu); unused by the static tags */
	ccl_context* ctx = nullptr;
	std::string source;
	void* module = nullptr;
	std::vector<ccl_kernel*> kernels;
};

struct ccl_prof {
	std::vector<ccl_queue*> queues;
	unsigned long long duration_ns;
};

/* ---- error plumbing ---- */

/* true (and *err set, CLO_ERROR_LIBRARY) when `e` is a CUDA failure */
bool clo_cuda_failed(cudaError_t e, GError** err, const char* what);

#define CLO_CUDA_TRY(call, err, label) \
	do { if (clo_cuda_failed((call), (err), #call)) goto label; } while (0)

/* ---- queue helpers ---- */

/* RAII device switch to the queue's device */
struct CloDeviceGuard {
	int prev;
	explicit CloDeviceGuard(int ordinal) {
		prev = -1;
		cudaGetDevice(&prev);
		if (prev != ordinal) cudaSetDevice(ordinal);
		else prev = -1;
	}
	~CloDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

/* begin / end an operation on a queue: returns the event that will mark its end */
ccl_event* clo_queue_begin(ccl_queue* cq, const char* name);
void clo_queue_end(ccl_queue* cq, ccl_event* evt);

/* handle registry (live-handle set: makes double destroy harmless and backs
 * ccl_wrapper_memcheck) */
void clo_handle_add(void* h);
bool clo_handle_alive(void* h);
bool clo_handle_remove(void* h);

/* launch counter (clo_b200_launch_count) */
extern std::atomic<unsigned long long> clo_launches;
#define CLO_COUNT_LAUNCH(n) (clo_launches.fetch_add((n), std::memory_order_relaxed))

/* cached device properties */
int clo_sm_count(int ordinal);

/* scratch buffer that only grows; freed by its owner */
struct CloScratch {
	void* ptr = nullptr;
	size_t size = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= size) return cudaSuccess;
		if (ptr) cudaFree(ptr);
		ptr = nullptr; size = 0;
		cudaError_t e = cudaMalloc(&ptr, bytes);
		if (e == cudaSuccess) size = bytes;
		return e;
	}
	void release() { if (ptr) cudaFree(ptr); ptr = nullptr; size = 0; }
};

/* stable multi-way partition by (key, global index) splitters (partition.cu) */
cudaError_t clo_partition_v2(CloScratch& work, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		void* keys_out, uint32_t* payload_out, size_t n, uint64_t gidx0, const void* splitter_keys,
		const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out, int sm_count, cudaStream_t stream,
		const char** err_msg);

cudaError_t clo_partition_count_stage(CloScratch& work, size_t elem_size, const void* keys_in, size_t n, uint64_t gidx0,
		const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out, int sm_count,
		cudaStream_t stream, const char** err_msg);
cudaError_t clo_partition_scatter_stage(CloScratch& work, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		size_t n, uint64_t gidx0, const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts,
		const uint64_t* first_slot, void* const* dests, void* const* vdests, const int* ok, int sm_count,
		cudaStream_t stream, const char** err_msg);

/* hashed 64-bit seeds for a CLO_RNG_HASH string outside the menu, built with NVRTC (jit.cu) */
cudaError_t clo_jit_seed_hash(const char* hash, unsigned long long* seeds_dev, size_t count, unsigned long long gid0,
		unsigned long long main_seed, cudaStream_t stream, std::string& msg);

#endif
