// tools/ubench_match.cu -- microbenchmarks that decide how the onesweep ranks keys:
//  (a) throughput of an 8-bit digit match built from 8 ballots,
//  (b) throughput of the hardware match.any.sync,
//  (c) whether same-address shared-memory atomics issued by one warp instruction are
//      applied in lane order (undocumented; measured only, never relied upon).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_match tools/ubench_match.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned match_ballot(unsigned d) {
	unsigned peers = 0xffffffffu;
#pragma unroll
	for (int b = 0; b < 8; ++b) {
		asm("{\n\t.reg .pred p;\n\t.reg .b32 m, t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
			"vote.sync.ballot.b32 m, p, 0xffffffff;\n\t@!p not.b32 m, m;\n\tand.b32 %0, %0, m;\n\t}"
			: "+r"(peers) : "r"(d), "r"(1u << b));
	}
	return peers;
}

template <int MODE, int ILP>
__global__ void k_match(const unsigned* in, unsigned* out, int iters, long long* cycles) {
	unsigned x[ILP];
	for (int i = 0; i < ILP; ++i) x[i] = in[(threadIdx.x + blockIdx.x * blockDim.x) * ILP + i];
	unsigned acc = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			unsigned d = (x[i] >> ((it & 3) * 8)) & 255u;
			unsigned p = MODE == 0 ? match_ballot(d) : __match_any_sync(0xffffffffu, d);
			acc += __popc(p);
			x[i] = x[i] * 1664525u + 1013904223u;
		}
	}
	long long t1 = clock64();
	out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_atom_order(unsigned* violations, unsigned* samples, int iters, unsigned seed) {
	__shared__ unsigned cnt[8][256];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	unsigned s = seed + threadIdx.x * 7919u + blockIdx.x * 104729u;
	unsigned bad = 0, tot = 0;
	for (int it = 0; it < iters; ++it) {
		for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
		__syncwarp();
		s = s * 1664525u + 1013904223u;
		unsigned d = (s >> 24) & ((it & 1) ? 255u : 15u);   // alternate sparse / dense collisions
		unsigned old = atomicAdd(&cnt[warp][d], 1u);
		unsigned peers = __match_any_sync(0xffffffffu, d);
		unsigned want = __popc(peers & ((1u << lane) - 1));
		bad += (old != want);
		tot += 1;
		__syncwarp();
	}
	atomicAdd(violations, bad);
	atomicAdd(samples, tot);
}

template <int MODE, int ILP>
void run(const char* name, const unsigned* d_in, unsigned* d_out, long long* d_cyc, int warps_per_sm_target) {
	int iters = 2000;
	int blocks = 148 * 4, threads = 256;  // 32 warps / SM
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_match<MODE, ILP><<<blocks, threads>>>(d_in, d_out, 10, d_cyc);
	cudaEventRecord(a);
	k_match<MODE, ILP><<<blocks, threads>>>(d_in, d_out, iters, d_cyc);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
	double matches = (double) blocks * threads / 32 * iters * ILP;  // warp-level matches
	printf("%-28s ILP=%d  %.3f ms  %.1f G warp-matches/s  (%.2f SM-cycles per warp-match at 1 warp; block0 cycles/match %.1f)\n",
		name, ILP, ms, matches / ms / 1e6, 0.0, (double) cyc / iters / ILP);
	(void) warps_per_sm_target;
}

int main() {
	unsigned *d_in, *d_out; long long* d_cyc;
	size_t n = 148 * 4 * 256 * 8;
	unsigned* h = (unsigned*) malloc(n * 4);
	for (size_t i = 0; i < n; ++i) h[i] = (unsigned) rand() * 2654435761u;
	cudaMalloc(&d_in, n * 4); cudaMalloc(&d_out, n * 4); cudaMalloc(&d_cyc, 8);
	cudaMemcpy(d_in, h, n * 4, cudaMemcpyHostToDevice);
	run<0, 1>("ballot x8", d_in, d_out, d_cyc, 32);
	run<0, 4>("ballot x8", d_in, d_out, d_cyc, 32);
	run<1, 1>("match.any.sync", d_in, d_out, d_cyc, 32);
	run<1, 4>("match.any.sync", d_in, d_out, d_cyc, 32);
	run<1, 8>("match.any.sync", d_in, d_out, d_cyc, 32);
	unsigned *d_v; cudaMalloc(&d_v, 8); cudaMemset(d_v, 0, 8);
	k_atom_order<<<148 * 2, 256>>>(d_v, d_v + 1, 4000, 12345u);
	unsigned v[2]; cudaMemcpy(v, d_v, 8, cudaMemcpyDeviceToHost);
	printf("same-address ATOMS lane-order violations: %u of %u samples\n", v[0], v[1]);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
