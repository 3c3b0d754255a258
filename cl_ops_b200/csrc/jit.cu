/*
 * jit.cu -- arbitrary `compare` / `get_key` macro strings for the comparison sorts.
 *
 * The reference splices the caller's OpenCL C macro bodies into its kernels and builds them at
 * run time (/root/reference/src/cl_ops/sort/clo_sort_abstract.c:144-168: CLO_SORT_ELEM_TYPE,
 * CLO_SORT_KEY_TYPE, CLO_SORT_COMPARE(a,b), CLO_SORT_KEY_GET(x)).  The precompiled kernels of
 * this library cover a fixed menu of strings (sort.cu); anything else takes this path: the same
 * macros are spliced into a CUDA C source of the canonical bitonic network
 * (clo_sort_sbitonic.cl:38-69, one compare-exchange step per launch, exactly the reference's
 * host loop clo_sort_sbitonic.c:73-118) and of the gselect rank kernel (clo_sort_gselect.cl:38-57)
 * and compiled with NVRTC for sm_100a.  The macro bodies are C expressions; OpenCL's scalar
 * type names (uchar, ushort, uint, ulong) are provided as typedefs.
 *
 * NVRTC and the CUDA driver are bound at run time (dlopen / cudaGetDriverEntryPoint), so the
 * library itself links against neither and still loads on a machine without a GPU.
 */
#include "clo_internal.h"
#include "sort_common.h"

#include <dlfcn.h>
#include <cstdarg>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

/* ---- NVRTC, bound lazily */
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
	void* h = nullptr;
	nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
	nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
	nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
	nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
	nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
	nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
	nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
	bool ok = false;
};

Nvrtc& nvrtc() {
	static Nvrtc n;
	if (n.h) return n;
	const char* names[] = { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so" };
	for (const char* nm : names) if ((n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
	if (!n.h) return n;
#define CLO_BIND(f) *(void**) &n.f = dlsym(n.h, "nvrtc" #f)
	CLO_BIND(CreateProgram); CLO_BIND(CompileProgram); CLO_BIND(GetProgramLogSize); CLO_BIND(GetProgramLog);
	CLO_BIND(GetCUBINSize); CLO_BIND(GetCUBIN); CLO_BIND(DestroyProgram);
#undef CLO_BIND
	n.ok = n.CreateProgram && n.CompileProgram && n.GetProgramLogSize && n.GetProgramLog && n.GetCUBINSize && n.GetCUBIN && n.DestroyProgram;
	return n;
}

/* ---- the few driver entry points needed to load and launch the compiled module */
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
struct Driver {
	int (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
	int (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
	int (*ModuleUnload)(CUmodule) = nullptr;
	int (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, cudaStream_t, void**, void**) = nullptr;
	bool ok = false;
};

Driver& driver() {
	static Driver d;
	static bool tried = false;
	if (tried) return d;
	tried = true;
	cudaDriverEntryPointQueryResult qr;
	auto get = [&](const char* name, void** fn) {
		return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) == cudaSuccess && *fn != nullptr;
	};
	d.ok = get("cuModuleLoadData", (void**) &d.ModuleLoadData) && get("cuModuleGetFunction", (void**) &d.ModuleGetFunction) &&
		get("cuModuleUnload", (void**) &d.ModuleUnload) && get("cuLaunchKernel", (void**) &d.LaunchKernel);
	return d;
}

const char* c_type_of(CloType t) {
	switch (t) {
	case CLO_CHAR: return "signed char";
	case CLO_UCHAR: return "unsigned char";
	case CLO_SHORT: return "short";
	case CLO_USHORT: return "unsigned short";
	case CLO_INT: return "int";
	case CLO_UINT: return "unsigned int";
	case CLO_LONG: return "long long";
	case CLO_ULONG: return "unsigned long long";
	case CLO_FLOAT: return "float";
	case CLO_DOUBLE: return "double";
	default: return nullptr;
	}
}

const char kJitBody[] = R"SRC(
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long long ulong;

/* "a must come after b"; padding elements (N not a power of two) come after everything */
__device__ __forceinline__ bool clo_must_swap(CLO_SORT_ELEM_TYPE a, CLO_SORT_ELEM_TYPE b, unsigned char pa, unsigned char pb, int padded) {
	CLO_SORT_KEY_TYPE ka = (CLO_SORT_KEY_TYPE) (CLO_SORT_KEY_GET(a));
	CLO_SORT_KEY_TYPE kb = (CLO_SORT_KEY_TYPE) (CLO_SORT_KEY_GET(b));
	const bool c = (CLO_SORT_COMPARE(ka, kb)) ? true : false;
	if (padded) return (pa > pb) || (pa == pb && c);
	return c;
}

/* one compare-exchange step of the canonical bitonic network */
extern "C" __global__ void clo_jit_bitonic_step(CLO_SORT_ELEM_TYPE* data, unsigned char* pad, unsigned long long npairs,
		int stage, int step, int padded) {
	const unsigned long long p = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= npairs) return;
	const unsigned long long stride = 1ull << (step - 1);
	const unsigned long long i1 = p + (p / stride) * stride;
	const unsigned long long i2 = i1 + stride;
	const bool desc = (p >> (stage - 1)) & 1;
	const CLO_SORT_ELEM_TYPE a = data[i1], b = data[i2];
	const unsigned char pa = padded ? pad[i1] : 0, pb = padded ? pad[i2] : 0;
	if (clo_must_swap(a, b, pa, pb, padded) != desc) {
		data[i1] = b; data[i2] = a;
		if (padded) { pad[i1] = pb; pad[i2] = pa; }
	}
}

extern "C" __global__ void clo_jit_init_pad(unsigned char* pad, unsigned long long n, unsigned long long np2) {
	const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < np2) pad[i] = i >= n ? 1 : 0;
}

/* gselect: rank = #{ i : COMPARE(key_gid, key_i) || (key_i == key_gid && i < gid) } */
extern "C" __global__ void clo_jit_gselect(const CLO_SORT_ELEM_TYPE* in, CLO_SORT_ELEM_TYPE* out, unsigned long long n) {
	const unsigned long long gid = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (gid >= n) return;
	const CLO_SORT_ELEM_TYPE mine = in[gid];
	const CLO_SORT_KEY_TYPE kg = (CLO_SORT_KEY_TYPE) (CLO_SORT_KEY_GET(mine));
	unsigned long long rank = 0;
	for (unsigned long long i = 0; i < n; ++i) {
		const CLO_SORT_ELEM_TYPE e = in[i];
		const CLO_SORT_KEY_TYPE ki = (CLO_SORT_KEY_TYPE) (CLO_SORT_KEY_GET(e));
		if ((CLO_SORT_COMPARE(kg, ki)) || (ki == kg && i < gid)) ++rank;
	}
	out[rank] = mine;
}

/* satradix with an arbitrary get_key: the key of every element, as raw bits zero-extended to a
 * 32- or 64-bit word (clo_sort_satradix.cl:58-61 takes bit b of `CLO_SORT_KEY_GET(value)`), next
 * to the element's index */
extern "C" __global__ void clo_jit_extract_keys(const CLO_SORT_ELEM_TYPE* in, void* keys, unsigned int* idx, unsigned long long n) {
	const unsigned long long i = (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const CLO_SORT_ELEM_TYPE x = in[i];
	const CLO_SORT_KEY_TYPE k = (CLO_SORT_KEY_TYPE) (CLO_SORT_KEY_GET(x));
	unsigned long long raw;
	if (sizeof(k) == 1) raw = *reinterpret_cast<const unsigned char*>(&k);
	else if (sizeof(k) == 2) raw = *reinterpret_cast<const unsigned short*>(&k);
	else if (sizeof(k) == 4) raw = *reinterpret_cast<const unsigned int*>(&k);
	else raw = *reinterpret_cast<const unsigned long long*>(&k);
	if (sizeof(k) <= 4) reinterpret_cast<unsigned int*>(keys)[i] = (unsigned int) raw;
	else reinterpret_cast<unsigned long long*>(keys)[i] = raw;
	idx[i] = (unsigned int) i;
}
)SRC";

} // namespace

struct CloJitSort {
	CUmodule mod = nullptr;
	CUfunction f_step = nullptr, f_pad = nullptr, f_gselect = nullptr, f_extract = nullptr;
	std::string source;
	CloScratch pad;
};

/* Compile the network for (elem, key, compare, get_key).  Returns NULL and fills `msg` (the
 * compiler log on a syntax error) on failure. */
CloJitSort* clo_jit_sort_new(CloType elem_type, CloType key_type, const char* compare, const char* get_key, std::string& msg) {
	const char* et = c_type_of(elem_type);
	const char* kt = c_type_of(key_type);
	if (!et || !kt) { msg = "unsupported element or key type for a run-time compiled sort"; return nullptr; }
	Nvrtc& rt = nvrtc();
	if (!rt.ok) { msg = "custom compare/get_key strings need NVRTC (libnvrtc.so.12 not found)"; return nullptr; }
	Driver& drv = driver();
	if (!drv.ok) { msg = "custom compare/get_key strings need the CUDA driver (cuModuleLoadData not available)"; return nullptr; }
	std::string src;
	src += "#define CLO_SORT_ELEM_TYPE "; src += et; src += "\n";
	src += "#define CLO_SORT_KEY_TYPE "; src += kt; src += "\n";
	src += "#define CLO_SORT_COMPARE(a, b) "; src += (compare && *compare) ? compare : "((a) > (b))"; src += "\n";
	src += "#define CLO_SORT_KEY_GET(x) "; src += (get_key && *get_key) ? get_key : "(x)"; src += "\n";
	src += kJitBody;
	nvrtcProgram prog = nullptr;
	if (rt.CreateProgram(&prog, src.c_str(), "clo_sort_jit.cu", 0, nullptr, nullptr) != 0) { msg = "nvrtcCreateProgram failed"; return nullptr; }
	const char* opts[] = { "--gpu-architecture=sm_100a", "--std=c++17", "-default-device" };
	const nvrtcResult rc = rt.CompileProgram(prog, 3, opts);
	if (rc != 0) {
		size_t ls = 0;
		rt.GetProgramLogSize(prog, &ls);
		std::vector<char> log(ls + 1, 0);
		if (ls) rt.GetProgramLog(prog, log.data());
		msg = "the compare / get_key strings do not compile: ";
		msg += log.data();
		rt.DestroyProgram(&prog);
		return nullptr;
	}
	size_t cs = 0;
	rt.GetCUBINSize(prog, &cs);
	std::vector<char> cubin(cs);
	rt.GetCUBIN(prog, cubin.data());
	rt.DestroyProgram(&prog);
	CloJitSort* j = new CloJitSort();
	j->source = src;
	cudaFree(0);      /* make sure the primary context exists before the driver call */
	if (drv.ModuleLoadData(&j->mod, cubin.data()) != 0 ||
			drv.ModuleGetFunction(&j->f_step, j->mod, "clo_jit_bitonic_step") != 0 ||
			drv.ModuleGetFunction(&j->f_pad, j->mod, "clo_jit_init_pad") != 0 ||
			drv.ModuleGetFunction(&j->f_gselect, j->mod, "clo_jit_gselect") != 0 ||
			drv.ModuleGetFunction(&j->f_extract, j->mod, "clo_jit_extract_keys") != 0) {
		msg = "loading the run-time compiled sort module failed";
		if (j->mod) drv.ModuleUnload(j->mod);
		delete j;
		return nullptr;
	}
	return j;
}

void clo_jit_sort_free(CloJitSort* j) {
	if (!j) return;
	if (j->mod) driver().ModuleUnload(j->mod);
	j->pad.release();
	delete j;
}

const char* clo_jit_sort_source(CloJitSort* j) { return j ? j->source.c_str() : nullptr; }

/* in-place bitonic sort of n elements (any n: padded to the next power of two with flags) */
cudaError_t clo_jit_bitonic_sort(CloJitSort* j, size_t elem_size, void* data, size_t n, cudaStream_t stream) {
	if (n < 2) return cudaSuccess;
	Driver& drv = driver();
	size_t np2 = 1; int log = 0;
	while (np2 < n) { np2 <<= 1; ++log; }
	int padded = np2 != n;
	void* work = data;
	unsigned char* pad = nullptr;
	cudaError_t e;
	CloScratch tmp;
	if (padded) {
		/* the network runs on a padded copy: [np2 elements][np2 pad flags] */
		if ((e = j->pad.reserve(np2 * elem_size + np2)) != cudaSuccess) return e;
		work = j->pad.ptr;
		pad = (unsigned char*) j->pad.ptr + np2 * elem_size;
		if ((e = cudaMemsetAsync(work, 0, np2 * elem_size, stream)) != cudaSuccess) return e;
		if ((e = cudaMemcpyAsync(work, data, n * elem_size, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
		unsigned long long nn = n, pp = np2;
		void* args[] = { &pad, &nn, &pp };
		if (drv.LaunchKernel(j->f_pad, (unsigned) ((np2 + 255) / 256), 1, 1, 256, 1, 1, 0, stream, args, nullptr) != 0) return cudaErrorLaunchFailure;
		CLO_COUNT_LAUNCH(1);
	}
	unsigned long long npairs = np2 / 2;
	for (int stage = 1; stage <= log; ++stage)
		for (int step = stage; step >= 1; --step) {
			void* args[] = { &work, &pad, &npairs, &stage, &step, &padded };
			if (drv.LaunchKernel(j->f_step, (unsigned) ((npairs + 255) / 256), 1, 1, 256, 1, 1, 0, stream, args, nullptr) != 0) return cudaErrorLaunchFailure;
			CLO_COUNT_LAUNCH(1);
		}
	if (padded) return cudaMemcpyAsync(data, work, n * elem_size, cudaMemcpyDeviceToDevice, stream);
	return cudaSuccess;
}

cudaError_t clo_jit_gselect_sort(CloJitSort* j, const void* in, void* out, size_t n, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	unsigned long long nn = n;
	void* args[] = { (void*) &in, &out, &nn };
	if (driver().LaunchKernel(j->f_gselect, (unsigned) ((n + 255) / 256), 1, 1, 256, 1, 1, 0, stream, args, nullptr) != 0) return cudaErrorLaunchFailure;
	CLO_COUNT_LAUNCH(1);
	return cudaSuccess;
}

/* satradix for a get_key string outside the menu: extract (key, index) with the run-time compiled
 * kernel, sort the pairs by the key's raw bits with the library's own stable LSD passes, gather
 * the elements.  Same result as the reference's passes over CLO_SORT_KEY_GET(value): a stable
 * sort by raw key bits, ascending. */
namespace {
template <typename T>
__global__ void clo_radix_gather(const T* __restrict__ in, const unsigned int* __restrict__ idx, T* __restrict__ out, size_t n) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = in[idx[i]];
}
}

cudaError_t clo_jit_radix_sort(CloJitSort* j, CloRadixState* rs, int sm_count, size_t elem_size, size_t key_size,
		uint32_t sorted_bits, const void* in, void* out, size_t n, cudaStream_t stream, const char** err_msg) {
	if (n == 0) return cudaSuccess;
	if (n > 0xffffffffull) { if (err_msg) *err_msg = "satradix with a run-time compiled get_key sorts at most 2^32 elements"; return cudaErrorInvalidValue; }
	const size_t kw = key_size <= 4 ? 4 : 8;
	const size_t keys_bytes = (n * kw + 255) & ~(size_t) 255, idx_bytes = (n * 4 + 255) & ~(size_t) 255;
	const bool in_place = in == out;
	cudaError_t e;
	if ((e = j->pad.reserve(keys_bytes + idx_bytes + (in_place ? n * elem_size : 0))) != cudaSuccess) return e;
	void* keys = j->pad.ptr;
	unsigned int* idx = (unsigned int*) ((char*) j->pad.ptr + keys_bytes);
	const void* src = in;
	if (in_place) {
		void* copy = (char*) j->pad.ptr + keys_bytes + idx_bytes;
		if ((e = cudaMemcpyAsync(copy, in, n * elem_size, cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
		src = copy;
	}
	unsigned long long nn = n;
	void* args[] = { (void*) &src, &keys, &idx, &nn };
	if (driver().LaunchKernel(j->f_extract, (unsigned) ((n + 255) / 256), 1, 1, 256, 1, 1, 0, stream, args, nullptr) != 0) return cudaErrorLaunchFailure;
	CLO_COUNT_LAUNCH(1);
	CloKeySpec ks = {};
	ks.mask = ~0ull; ks.shift = 0; ks.elem_bits = ks.key_bits = (uint32_t) (8 * kw);
	ks.elem_signed = 0; ks.key_kind = CLO_KIND_UNSIGNED; ks.descending = 0; ks.identity = 1;
	if ((e = clo_radix_sort(rs, sm_count, kw, ks, sorted_bits < 8 * key_size ? sorted_bits : (uint32_t) (8 * key_size), keys, keys, idx, idx, n, stream, err_msg)) != cudaSuccess) return e;
	const unsigned grid = (unsigned) ((n + 255) / 256);
	switch (elem_size) {
	case 1: clo_radix_gather<unsigned char><<<grid, 256, 0, stream>>>((const unsigned char*) src, idx, (unsigned char*) out, n); break;
	case 2: clo_radix_gather<unsigned short><<<grid, 256, 0, stream>>>((const unsigned short*) src, idx, (unsigned short*) out, n); break;
	case 4: clo_radix_gather<unsigned int><<<grid, 256, 0, stream>>>((const unsigned int*) src, idx, (unsigned int*) out, n); break;
	default: clo_radix_gather<unsigned long long><<<grid, 256, 0, stream>>>((const unsigned long long*) src, idx, (unsigned long long*) out, n); break;
	}
	CLO_COUNT_LAUNCH(1);
	return cudaGetLastError();
}

/* =====================================================================================
 * cf4ocl2 program / kernel objects on NVRTC: what the reference's clo_rng_bench.c:176-312 and
 * tests/test_rng.c:85-118 do with the source string of clo_rng_get_source() -- concatenate a
 * small OpenCL C kernel, build it, set its arguments, enqueue it.  The kernel text is OpenCL C;
 * the prelude below maps the handful of OpenCL spellings those kernels use onto CUDA C.
 * ===================================================================================== */
namespace {

const char kOclPrelude[] = R"SRC(
#define __kernel extern "C" __global__
#define __global
#define __constant const
#define __private
#define __local __shared__
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long long ulong;
#ifndef UINT_MAX
#define UINT_MAX 0xffffffffu
#endif
#ifndef INT_MAX
#define INT_MAX 2147483647
#endif
__device__ __forceinline__ unsigned long long get_global_id(int d) { return d == 0 ? (unsigned long long) blockIdx.x * blockDim.x + threadIdx.x : 0; }
__device__ __forceinline__ unsigned long long get_local_id(int d) { return d == 0 ? threadIdx.x : 0; }
__device__ __forceinline__ unsigned long long get_group_id(int d) { return d == 0 ? blockIdx.x : 0; }
__device__ __forceinline__ unsigned long long get_local_size(int d) { return d == 0 ? blockDim.x : 1; }
__device__ __forceinline__ unsigned long long get_global_size(int d) { return d == 0 ? (unsigned long long) gridDim.x * blockDim.x : 1; }
)SRC";

struct ArgSlot { std::vector<unsigned char> bytes; };

} // namespace

struct ccl_arg { unsigned magic; std::vector<unsigned char> bytes; };
struct ccl_kernel {
	ccl_program* prg;
	CUfunction fn;
	std::string name;
	std::vector<ArgSlot> args;
};

extern "C" CCLProgram* ccl_program_new_from_source(CCLContext* ctx, const char* src, GError** err) {
	if (!ctx || !src) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ccl_program_new_from_source: NULL argument"); return NULL; }
	ccl_program* p = new ccl_program();
	p->tag = "run-time compiled (NVRTC, sm_100a)";
	p->ctx = ctx;
	p->source = src;
	ccl_context_ref(ctx);
	clo_handle_add(p);
	return p;
}

extern "C" cl_bool ccl_program_build(CCLProgram* prg, const char* options, GError** err) {
	if (!prg || !prg->ctx) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ccl_program_build: not a source program"); return CL_FALSE; }
	Nvrtc& rt = nvrtc();
	Driver& drv = driver();
	if (!rt.ok || !drv.ok) { g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "building a program needs NVRTC and the CUDA driver"); return CL_FALSE; }
	/* OpenCL build options: only the -D definitions have a meaning here */
	std::vector<std::string> defs;
	if (options) {
		std::string o(options);
		size_t i = 0;
		while ((i = o.find("-D", i)) != std::string::npos) {
			i += 2;
			while (i < o.size() && o[i] == ' ') ++i;
			size_t e = i;
			while (e < o.size() && o[e] != ' ') ++e;
			if (e > i) defs.push_back("-D" + o.substr(i, e - i));
			i = e;
		}
	}
	std::vector<const char*> opts = { "--gpu-architecture=sm_100a", "--std=c++17", "-default-device" };
	for (const std::string& d : defs) opts.push_back(d.c_str());
	const std::string full = std::string(kOclPrelude) + prg->source;
	nvrtcProgram np = nullptr;
	if (rt.CreateProgram(&np, full.c_str(), "clo_program.cu", 0, nullptr, nullptr) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "nvrtcCreateProgram failed");
		return CL_FALSE;
	}
	if (rt.CompileProgram(np, (int) opts.size(), opts.data()) != 0) {
		size_t ls = 0;
		rt.GetProgramLogSize(np, &ls);
		std::vector<char> log(ls + 1, 0);
		if (ls) rt.GetProgramLog(np, log.data());
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "program build failed: %s", log.data());
		rt.DestroyProgram(&np);
		return CL_FALSE;
	}
	size_t cs = 0;
	rt.GetCUBINSize(np, &cs);
	std::vector<char> cubin(cs);
	rt.GetCUBIN(np, cubin.data());
	rt.DestroyProgram(&np);
	CloDeviceGuard g(prg->ctx->dev.ordinal);
	cudaFree(0);
	CUmodule mod = nullptr;
	if (drv.ModuleLoadData(&mod, cubin.data()) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "loading the built program failed");
		return CL_FALSE;
	}
	prg->module = mod;
	return CL_TRUE;
}

extern "C" CCLKernel* ccl_program_get_kernel(CCLProgram* prg, const char* name, GError** err) {
	if (!prg || !prg->module || !name) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ccl_program_get_kernel: program not built"); return NULL; }
	for (ccl_kernel* k : prg->kernels) if (k->name == name) return k;
	CUfunction fn = nullptr;
	if (driver().ModuleGetFunction(&fn, (CUmodule) prg->module, name) != 0) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel '%s' not found in program", name);
		return NULL;
	}
	ccl_kernel* k = new ccl_kernel();
	k->prg = prg; k->fn = fn; k->name = name;
	prg->kernels.push_back(k);        /* owned by the program, as in cf4ocl */
	return k;
}

extern "C" void ccl_program_destroy(CCLProgram* prg) {
	if (!prg || !prg->ctx || !clo_handle_alive(prg)) return;
	clo_handle_remove(prg);
	for (ccl_kernel* k : prg->kernels) delete k;
	if (prg->module) {
		CloDeviceGuard g(prg->ctx->dev.ordinal);
		driver().ModuleUnload((CUmodule) prg->module);
	}
	ccl_context_unref(prg->ctx);
	delete prg;
}

namespace {
std::mutex g_arg_mtx;
std::vector<ccl_arg*>& live_args() { static std::vector<ccl_arg*> v; return v; }

void set_args_va(ccl_kernel* k, va_list ap) {
	k->args.clear();
	for (void* a = va_arg(ap, void*); a; a = va_arg(ap, void*)) {
		ArgSlot s;
		if (clo_handle_alive(a)) {                       /* a CCLBuffer: its device address */
			void* dptr = ((ccl_buffer*) a)->ptr;
			s.bytes.assign((unsigned char*) &dptr, (unsigned char*) &dptr + sizeof(void*));
		} else {                                         /* a private value from ccl_arg_priv / ccl_arg_new */
			std::lock_guard<std::mutex> lk(g_arg_mtx);
			std::vector<ccl_arg*>& v = live_args();
			bool found = false;
			for (size_t i = 0; i < v.size(); ++i) if (v[i] == a) { found = true; v.erase(v.begin() + i); break; }
			if (!found) continue;
			s.bytes = ((ccl_arg*) a)->bytes;
			delete (ccl_arg*) a;
		}
		k->args.push_back(s);
	}
}

CCLEvent* enqueue(ccl_kernel* k, CCLQueue* cq, cl_uint dims, const size_t* gws, const size_t* lws, GError** err) {
	if (!k || !cq || dims != 1 || !gws) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ccl_kernel_enqueue_ndrange: one-dimensional ranges only"); return NULL; }
	size_t l = (lws && *lws) ? *lws : 0;
	if (!l) { l = 256; while (l > 1 && (*gws % l)) l >>= 1; }
	if (*gws % l) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "global work size %zu is not a multiple of the local size %zu", *gws, l); return NULL; }
	std::vector<void*> ptrs;
	for (ArgSlot& s : k->args) ptrs.push_back(s.bytes.data());
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq, k->name.c_str());
	const int rc = *gws ? driver().LaunchKernel(k->fn, (unsigned) (*gws / l), 1, 1, (unsigned) l, 1, 1, 0, cq->stream, ptrs.data(), nullptr) : 0;
	clo_queue_end(cq, evt);
	CLO_COUNT_LAUNCH(1);
	if (rc != 0) { g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "cuLaunchKernel failed (%d) for '%s'", rc, k->name.c_str()); return NULL; }
	return evt;
}
} // namespace

extern "C" CCLArg* ccl_arg_new(void* value, size_t size) {
	ccl_arg* a = new ccl_arg();
	a->magic = 0xC10A26u;
	a->bytes.assign((unsigned char*) value, (unsigned char*) value + size);
	std::lock_guard<std::mutex> lk(g_arg_mtx);
	live_args().push_back(a);
	return a;
}

extern "C" void ccl_kernel_set_args(CCLKernel* krnl, ...) {
	if (!krnl) return;
	va_list ap;
	va_start(ap, krnl);
	set_args_va(krnl, ap);
	va_end(ap);
}

extern "C" CCLEvent* ccl_kernel_enqueue_ndrange(CCLKernel* krnl, CCLQueue* cq, cl_uint work_dim, const size_t* gwo,
		const size_t* gws, const size_t* lws, CCLEventWaitList* ewl, GError** err) {
	(void) gwo;
	if (ewl) ccl_event_wait_list_clear(ewl);      /* same-queue ordering is implied by the stream */
	return enqueue(krnl, cq, work_dim, gws, lws, err);
}

extern "C" CCLEvent* ccl_kernel_set_args_and_enqueue_ndrange(CCLKernel* krnl, CCLQueue* cq, cl_uint work_dim,
		const size_t* gwo, const size_t* gws, const size_t* lws, CCLEventWaitList* ewl, GError** err, ...) {
	(void) gwo;
	if (!krnl) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL kernel"); return NULL; }
	va_list ap;
	va_start(ap, err);
	set_args_va(krnl, ap);
	va_end(ap);
	if (ewl) ccl_event_wait_list_clear(ewl);
	return enqueue(krnl, cq, work_dim, gws, lws, err);
}

/* lws in: the caller's maximum (0 = none); out: a power of two <= 256.  With gws == NULL the
 * local size is made a divisor of real_ws, otherwise gws = real_ws rounded up to it. */
extern "C" cl_bool ccl_kernel_suggest_worksizes(CCLKernel* krnl, CCLDevice* dev, cl_uint dims, const size_t* real_ws,
		size_t* gws, size_t* lws, GError** err) {
	(void) krnl; (void) dev;
	if (dims != 1 || !real_ws || !lws) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "ccl_kernel_suggest_worksizes: one dimension only"); return CL_FALSE; }
	size_t cap = (*lws && *lws < 256) ? *lws : 256;
	size_t l = 1;
	while (l * 2 <= cap) l *= 2;
	while (l > 1 && l > *real_ws) l >>= 1;
	if (gws) {
		*gws = (*real_ws + l - 1) / l * l;
	} else {
		while (l > 1 && (*real_ws % l)) l >>= 1;
	}
	*lws = l;
	return CL_TRUE;
}

/* =====================================================================================
 * Seed hashes outside the menu: the reference defines CLO_RNG_HASH(x) as the caller's string
 * and builds it into its init kernel (clo_rng.c:101-109, clo_rng_init.cl:27-60).  Here the
 * string is built, once per distinct string, into a kernel that produces the hashed 64-bit
 * seeds; rng.cu turns them into generator states (clo_ulong2statetype).
 * ===================================================================================== */
namespace {
const char kHashBody[] = R"SRC(
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long long ulong;
#define KNUTH(x) x = ((x*2654435761) % 0x100000000)
#define XS1(x) \
	x = ((x >> 16) ^ x) * 0x45d9f3b; \
	x = ((x >> 16) ^ x) * 0x45d9f3b; \
	x = ((x >> 16) ^ x);
extern "C" __global__ void clo_jit_seed_hash(ulong* seeds, ulong count, ulong gid0, ulong main_seed) {
	const ulong i = (ulong) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	ulong seed = gid0 + i + main_seed;
	CLO_RNG_HASH(seed);
	seeds[i] = seed;
}
)SRC";

struct HashModule { CUmodule mod; CUfunction fn; };
std::mutex g_hash_mtx;
std::vector<std::pair<std::string, HashModule>>& hash_cache() { static std::vector<std::pair<std::string, HashModule>> v; return v; }
} // namespace

cudaError_t clo_jit_seed_hash(const char* hash, unsigned long long* seeds_dev, size_t count, unsigned long long gid0,
		unsigned long long main_seed, cudaStream_t stream, std::string& msg) {
	Nvrtc& rt = nvrtc();
	Driver& drv = driver();
	if (!rt.ok || !drv.ok) { msg = "a custom seed hash needs NVRTC and the CUDA driver"; return cudaErrorNotSupported; }
	HashModule hm = { nullptr, nullptr };
	{
		std::lock_guard<std::mutex> lk(g_hash_mtx);
		for (auto& e : hash_cache()) if (e.first == hash) hm = e.second;
	}
	if (!hm.fn) {
		std::string src = "#define CLO_RNG_HASH(x) ";
		src += hash; src += "\n"; src += kHashBody;
		nvrtcProgram np = nullptr;
		if (rt.CreateProgram(&np, src.c_str(), "clo_rng_hash.cu", 0, nullptr, nullptr) != 0) { msg = "nvrtcCreateProgram failed"; return cudaErrorUnknown; }
		const char* opts[] = { "--gpu-architecture=sm_100a", "--std=c++17", "-default-device" };
		if (rt.CompileProgram(np, 3, opts) != 0) {
			size_t ls = 0;
			rt.GetProgramLogSize(np, &ls);
			std::vector<char> log(ls + 1, 0);
			if (ls) rt.GetProgramLog(np, log.data());
			msg = "the seed hash does not compile: ";
			msg += log.data();
			rt.DestroyProgram(&np);
			return cudaErrorInvalidValue;
		}
		size_t cs = 0;
		rt.GetCUBINSize(np, &cs);
		std::vector<char> cubin(cs);
		rt.GetCUBIN(np, cubin.data());
		rt.DestroyProgram(&np);
		cudaFree(0);
		if (drv.ModuleLoadData(&hm.mod, cubin.data()) != 0 || drv.ModuleGetFunction(&hm.fn, hm.mod, "clo_jit_seed_hash") != 0) {
			msg = "loading the seed-hash module failed";
			return cudaErrorUnknown;
		}
		std::lock_guard<std::mutex> lk(g_hash_mtx);
		hash_cache().push_back({ hash, hm });
	}
	if (!count) return cudaSuccess;
	unsigned long long c = count;
	void* args[] = { &seeds_dev, &c, &gid0, &main_seed };
	if (drv.LaunchKernel(hm.fn, (unsigned) ((count + 255) / 256), 1, 1, 256, 1, 1, 0, stream, args, nullptr) != 0) {
		msg = "launching the seed-hash kernel failed";
		return cudaErrorLaunchFailure;
	}
	CLO_COUNT_LAUNCH(1);
	return cudaSuccess;
}
