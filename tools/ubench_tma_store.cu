// tools/ubench_tma_store.cu -- can the write-out of a staged radix tile leave the SM through the
// bulk-copy engine (cp.async.bulk shared -> global, SASS UBLKCP) instead of through LDS + STG?
// A tile of 8192 u32 keys holds 256 digit runs of ~32 keys; a run goes to an arbitrary 4-byte
// aligned global position.  cp.async.bulk needs 16-byte aligned addresses and sizes, so a run is
// staged at a shared-memory position congruent (mod 16 B) to its destination and leaves as
//   [masked 16 B head chunk] + [16 B-aligned body] + [masked 16 B tail chunk]
// (.cp_mask, sm_100+: UBLKCP.G.S.DST_G_BYTE_MASK).  This measures the engine's rate for such small
// copies and checks that the byte masks do not touch a neighbour's bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tma_store tools/ubench_tma_store.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

typedef unsigned int u32;
constexpr int RADIX = 256;


__device__ __forceinline__ void bulk_s2g(void* dst, u32 src, u32 bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g_mask(void* dst, u32 src, u32 mask) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group.cp_mask [%0], [%1], 16, %2;" :: "l"(dst), "r"(src), "h"((unsigned short) mask) : "memory");
}

// mode 0: head + body + tail (full run); mode 1: aligned body only
template <int MODE, bool FILL>
__global__ void __launch_bounds__(512, 2)
k_store(u32* __restrict__ out, const u32* __restrict__ G, const unsigned short* __restrict__ C,
		int num_tiles, int tile_keys) {
	extern __shared__ __align__(128) u32 smem[];
	u32* stage0 = smem;                      // [2][STAGE]
	const int STAGE = tile_keys + 6 * RADIX;
	__shared__ u32 s_S[RADIX], s_G[RADIX], s_scan[8];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	int b = 0;
	u32 g_next = 0, c_next = 0;
	if (tid < RADIX && blockIdx.x < num_tiles) { g_next = G[(size_t) blockIdx.x * RADIX + tid]; c_next = C[(size_t) blockIdx.x * RADIX + tid]; }
	for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, b ^= 1) {
		u32* stage = stage0 + b * STAGE;
		u32 g = 0, c = 0, S = 0;
		if (tid < RADIX) {
			g = g_next; c = c_next;
			if (t + gridDim.x < num_tiles) { g_next = G[(size_t) (t + gridDim.x) * RADIX + tid]; c_next = C[(size_t) (t + gridDim.x) * RADIX + tid]; }
			// region of 4-aligned size holding the run at offset g mod 4
			const u32 a = g & 3u;
			const u32 reg = c ? ((a + c + 3u) & ~3u) : 0u;
			u32 incl = reg;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const u32 v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
			if (lane == 31) s_scan[warp] = incl;
			asm volatile("bar.sync 1, 256;" ::: "memory");
			u32 off = 0;
			for (int w = 0; w < warp; ++w) off += s_scan[w];
			S = off + incl - reg + a;
			s_S[tid] = S; s_G[tid] = g;           // destination of staged index S
		}
		// the staging buffer of two tiles ago must have been read by the engine
		asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
		__syncthreads();
		if (FILL) {
			// every staged key carries the global index it must land on
			for (int d = warp; d < RADIX; d += 16) {
				const u32 cc = C[(size_t) t * RADIX + d], SS = s_S[d], gg = s_G[d];
				for (u32 r = lane; r < cc; r += 32) stage[SS + r] = gg + r;
			}
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			__syncthreads();
		}
		if (tid < RADIX && c) {
			const u32 sbase = (u32) __cvta_generic_to_shared(stage);
			u32 s = S, e = S + c;                 // staged [s, e)
			u32* gp = out + ((long long) g - (long long) S);   // gp + staged index = destination
			const u32 s_al = (s + 3u) & ~3u, e_al = e & ~3u;
			if (MODE == 0) {
				if (s_al > e_al) {
					// run inside one 16-byte chunk
					const u32 ch = s & ~3u;
					const u32 m = ((0xffffu << (4 * (s - ch))) & (0xffffu >> (4 * (ch + 4 - e)))) & 0xffffu;
					bulk_s2g_mask(gp + ch, sbase + ch * 4, m);
				} else {
					if (s < s_al) bulk_s2g_mask(gp + (s_al - 4), sbase + (s_al - 4) * 4, (0xffffu << (4 * (s - (s_al - 4)))) & 0xffffu);
					if (e_al > s_al) bulk_s2g(gp + s_al, sbase + s_al * 4, (e_al - s_al) * 4);
					if (e > e_al) bulk_s2g_mask(gp + e_al, sbase + e_al * 4, 0xffffu >> (4 * (e_al + 4 - e)));
				}
			} else {
				if (e_al > s_al) bulk_s2g(gp + s_al, sbase + s_al * 4, (e_al - s_al) * 4);
			}
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		} else {
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
	}
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
	const int log_n = argc > 1 ? atoi(argv[1]) : 28;
	const int tile = argc > 2 ? atoi(argv[2]) : 8192;
	const int STAGE = tile + 6 * RADIX;
	const int ctas_per_sm = tile > 8192 ? 1 : 2;
	const size_t n = (size_t) 1 << log_n;
	const int num_tiles = (int) (n / tile);
	// run lengths: multinomial(8192; 256 equal bins) per tile, like uniform random digits
	std::vector<unsigned short> C((size_t) num_tiles * RADIX);
	std::vector<u32> G((size_t) num_tiles * RADIX);
	std::vector<size_t> tot(RADIX, 0);
	u32 s = 12345u;
	for (int t = 0; t < num_tiles; ++t) {
		unsigned short* c = &C[(size_t) t * RADIX];
		for (int d = 0; d < RADIX; ++d) c[d] = 0;
		for (int k = 0; k < tile; ++k) { s = s * 1664525u + 1013904223u; c[s >> 24]++; }
		for (int d = 0; d < RADIX; ++d) tot[d] += c[d];
	}
	std::vector<size_t> base(RADIX, 0);
	for (int d = 1; d < RADIX; ++d) base[d] = base[d - 1] + tot[d - 1];
	for (int t = 0; t < num_tiles; ++t)
		for (int d = 0; d < RADIX; ++d) { G[(size_t) t * RADIX + d] = (u32) base[d]; base[d] += C[(size_t) t * RADIX + d]; }
	u32 *d_out, *d_G; unsigned short* d_C;
	CK(cudaMalloc(&d_out, n * 4 + 64)); CK(cudaMalloc(&d_G, G.size() * 4)); CK(cudaMalloc(&d_C, C.size() * 2));
	CK(cudaMemcpy(d_G, G.data(), G.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_C, C.data(), C.size() * 2, cudaMemcpyHostToDevice));
	int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	const size_t smem = (size_t) 2 * STAGE * 4;
	auto run = [&](auto kern, const char* name, bool check) {
		CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		CK(cudaMemset(d_out, 0xff, n * 4));
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		float best = 1e9f;
		for (int rep = 0; rep < (check ? 1 : 5); ++rep) {
			cudaEventRecord(e0);
			kern<<<ctas_per_sm * sms, 512, smem>>>(d_out, d_G, d_C, num_tiles, tile);
			cudaEventRecord(e1);
			CK(cudaEventSynchronize(e1));
			float ms; cudaEventElapsedTime(&ms, e0, e1);
			if (ms < best) best = ms;
		}
		CK(cudaGetLastError());
		printf("%-34s %8.3f ms  %7.1f GB/s written  %6.2f Mruns/s/SM\n", name, best, n * 4 / best * 1e-6,
			(double) num_tiles * RADIX / best * 1e-3 / sms);
		if (check) {
			std::vector<u32> h(n);
			CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
			size_t bad = 0;
			for (size_t i = 0; i < n; ++i) bad += (h[i] != (u32) i);
			printf("   check: %zu of %zu keys wrong%s\n", bad, n, bad ? "  <-- FAIL" : "  (byte masks exact)");
		}
	};
	run(k_store<0, true>, "fill + head/body/tail (checked)", true);
	run(k_store<0, false>, "head + body + tail", false);
	run(k_store<1, false>, "aligned body only", false);
	run(k_store<0, true>, "fill + head/body/tail", false);
	return 0;
}
