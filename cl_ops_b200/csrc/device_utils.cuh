/*
 * device_utils.cuh -- small device-side helpers shared by the kernels:
 * relaxed 64-bit descriptor loads/stores, streaming vector loads/stores,
 * warp scans.
 */
#ifndef CLO_DEVICE_UTILS_CUH
#define CLO_DEVICE_UTILS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

namespace clo {

typedef unsigned long long u64;
typedef unsigned int u32;

/* Descriptor words are single 64-bit words carrying flag and payload together,
 * so relaxed GPU-scope accesses are sufficient (no fence between flag and data). */
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
	u64 v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32 ld_relaxed(const u32* p) {
	u32 v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_relaxed(u32* p, u32 v) {
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

/* N elements of T moved as one naturally aligned access of N*sizeof(T) bytes */
template <typename T, int N>
struct alignas(sizeof(T) * N >= 16 ? 16 : sizeof(T) * N) Vec {
	T d[N];
};

template <int BYTES> struct RawVec;
template <> struct RawVec<1> { typedef unsigned char type; };
template <> struct RawVec<2> { typedef unsigned short type; };
template <> struct RawVec<4> { typedef unsigned int type; };
template <> struct RawVec<8> { typedef uint2 type; };
template <> struct RawVec<16> { typedef uint4 type; };

/* streaming (evict-first) load of a whole Vec<T,N>; sizeof(Vec) in {1,2,4,8,16} */
template <typename T, int N>
__device__ __forceinline__ void load_vec_cs(const T* p, T (&dst)[N]) {
	typedef typename RawVec<sizeof(T) * N>::type R;
	union { R r; T t[N]; } u;
	u.r = __ldcs(reinterpret_cast<const R*>(p));
#pragma unroll
	for (int i = 0; i < N; ++i) dst[i] = u.t[i];
}

template <typename T, int N>
__device__ __forceinline__ void store_vec_cs(T* p, const T (&src)[N]) {
	typedef typename RawVec<sizeof(T) * N>::type R;
	union { R r; T t[N]; } u;
#pragma unroll
	for (int i = 0; i < N; ++i) u.t[i] = src[i];
	__stcs(reinterpret_cast<R*>(p), u.r);
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v, int lane) {
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		T t = __shfl_up_sync(0xffffffffu, v, off);
		if (lane >= off) v += t;
	}
	return v;
}

template <typename T>
__device__ __forceinline__ T warp_reduce_sum(T v) {
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
	return v;
}

__device__ __forceinline__ u32 lanemask_lt() {
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

} // namespace clo

#endif
