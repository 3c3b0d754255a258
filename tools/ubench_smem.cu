// tools/ubench_smem.cu -- SM-level throughput of the shared-memory operations the onesweep
// pass is made of, with random 8-bit digits as addresses (32 warps resident per SM):
//   0 ATOMS.ADD with return, warp-private [256] table      (the rank)
//   1 RED (atomicAdd, result unused), warp-private table
//   2 LDS random word of a warp-private [256] table          (the base lookup)
//   3 STS random word in a 8192-word tile                    (the staging scatter)
//   4 LDS linear                                             (the write-out read)
//   5 ATOMS.ADD with return, packed u16 pairs, warp-private [128] words
//   6 ATOMS.ADD with return, CTA-shared [256] table
//   7 LDS.64 random (digit -> 8-byte entry)
//   8 SHFL.UP, 9 SHFL.UP + LDS linear (do shuffles share the shared-memory data pipe?)
// Output: SM cycles per warp instruction (lower bound on what a pass pays per 32 keys).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_smem tools/ubench_smem.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k_smem(unsigned* out, int iters, unsigned seed) {
	__shared__ __align__(16) unsigned tab[8192];
	const int warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < 8192; i += 256) tab[i] = i;
	__syncthreads();
	unsigned* wt = tab + warp * 256;
	unsigned x[ILP];
	for (int i = 0; i < ILP; ++i) x[i] = (seed + threadIdx.x * 7919u + blockIdx.x * 104729u + i * 31u) * 2654435761u;
	unsigned acc = 0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			const unsigned d = x[i] >> 24;
			if (MODE == 0) acc += atomicAdd(&wt[d], 1u);
			else if (MODE == 1) atomicAdd(&wt[d], 1u);
			else if (MODE == 2) acc += wt[d];
			else if (MODE == 3) tab[x[i] >> 19] = x[i];
			else if (MODE == 4) acc += tab[(threadIdx.x + 256 * i + it * 32) & 8191];
			else if (MODE == 5) acc += atomicAdd(&wt[d >> 1], (d & 1) ? 65536u : 1u);
			else if (MODE == 6) acc += atomicAdd(&tab[d], 1u);
			else if (MODE == 7) { uint2 v = reinterpret_cast<uint2*>(wt)[d & 127]; acc += v.x ^ v.y; }
			else if (MODE == 8) acc += __shfl_up_sync(0xffffffffu, x[i], 1);
			else if (MODE == 9) { acc += __shfl_up_sync(0xffffffffu, x[i], 1); acc += tab[(threadIdx.x + 256 * i + it * 32) & 8191]; }
			x[i] = x[i] * 1664525u + 1013904223u;
		}
	}
	out[threadIdx.x + blockIdx.x * blockDim.x] = acc + tab[threadIdx.x];
}

template <int MODE, int ILP>
void run(const char* name, unsigned* d_out, double sm_mhz) {
	const int iters = 4000, blocks = 148 * 4, threads = 256;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	k_smem<MODE, ILP><<<blocks, threads>>>(d_out, 10, 1u);
	cudaEventRecord(a);
	k_smem<MODE, ILP><<<blocks, threads>>>(d_out, iters, 1u);
	cudaEventRecord(b); cudaEventSynchronize(b);
	float ms; cudaEventElapsedTime(&ms, a, b);
	const double warp_instr_per_sm = 4.0 * threads / 32 * iters * ILP;
	printf("%-44s ILP=%d %.3f ms  %.2f SM-cycles per warp instruction (at %.0f MHz)\n", name, ILP, ms,
		ms * 1e-3 * sm_mhz * 1e6 / warp_instr_per_sm, sm_mhz);
}

int main() {
	unsigned* d_out; cudaMalloc(&d_out, 148 * 4 * 256 * 4);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double mhz = khz / 1000.0;
	run<0, 8>("ATOMS.ADD ret, warp-private, random digit", d_out, mhz);
	run<1, 8>("RED, warp-private, random digit", d_out, mhz);
	run<2, 8>("LDS random digit", d_out, mhz);
	run<3, 8>("STS random in tile", d_out, mhz);
	run<4, 8>("LDS linear", d_out, mhz);
	run<5, 8>("ATOMS.ADD ret packed u16", d_out, mhz);
	run<6, 8>("ATOMS.ADD ret, CTA-shared table", d_out, mhz);
	run<7, 8>("LDS.64 random", d_out, mhz);
	run<8, 8>("SHFL.UP", d_out, mhz);
	run<9, 8>("SHFL.UP + LDS linear (2 ops: same pipe if ~sum)", d_out, mhz);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
