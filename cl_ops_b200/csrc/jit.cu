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
)SRC";

} // namespace

struct CloJitSort {
	CUmodule mod = nullptr;
	CUfunction f_step = nullptr, f_pad = nullptr, f_gselect = nullptr;
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
			drv.ModuleGetFunction(&j->f_gselect, j->mod, "clo_jit_gselect") != 0) {
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
