// tools/ubench_atoms_dup.cu -- what do DUPLICATE addresses inside one warp instruction cost a
// shared-memory atomic on B200?  (skewed digits: Zipf keys, few distinct values, sorted input.)
// Groups of m consecutive lanes share one random digit; m = 1 (uniform keys) .. 32 (sorted keys).
//   0 atomicAdd(&row[d], 1) with return      (ptxas: ATOMS.POPC.INC)
//   1 atomicAdd(&row[d], 1), result unused   (RED)
//   2 packed u16 pair add with return        (ATOMS.ADD, the v6 placement)
//   3 packed u16 pair add, result unused
//   4 leader-only: run heads add the run length (what warp aggregation would leave)
// Output: SM cycles per warp instruction, 32 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_atoms_dup tools/ubench_atoms_dup.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, int m) {
	__shared__ __align__(16) unsigned tab[8 * 256];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int i = threadIdx.x; i < 8 * 256; i += 256) tab[i] = 0;
	__syncthreads();
	unsigned* wt = tab + warp * 256;
	unsigned x[ILP];
	// the same state in all lanes of a group -> the same digit
	for (int i = 0; i < ILP; ++i) x[i] = (1u + (unsigned) (lane / m) * 7919u + warp * 977u + blockIdx.x * 104729u + i * 31u) * 2654435761u;
	unsigned acc = 0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < ILP; ++i) {
			const unsigned d = x[i] >> 24;
			if (MODE == 0) acc += atomicAdd(&wt[d], 1u);
			else if (MODE == 1) atomicAdd(&wt[d], 1u);
			else if (MODE == 2) acc += atomicAdd(&wt[d >> 1], (d & 1) ? 65536u : 1u);
			else if (MODE == 3) atomicAdd(&wt[d >> 1], (d & 1) ? 65536u : 1u);
			else if (MODE == 4) { if (lane % m == 0) acc += atomicAdd(&wt[d >> 1], ((d & 1) ? 65536u : 1u) * m); }
			x[i] = x[i] * 1664525u + 1013904223u;
		}
	}
	out[threadIdx.x + blockIdx.x * blockDim.x] = acc + tab[threadIdx.x];
}

template <int MODE>
void run(const char* name, unsigned* d_out, double mhz) {
	printf("%-36s", name);
	for (int m = 1; m <= 32; m *= 2) {
		const int iters = 2000, blocks = 148 * 4, threads = 256, ILP = 8;
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		k<MODE, ILP><<<blocks, threads>>>(d_out, 10, m);
		cudaEventRecord(a);
		k<MODE, ILP><<<blocks, threads>>>(d_out, iters, m);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		printf("  m=%-2d %6.2f", m, ms * 1e-3 * mhz * 1e6 / (4.0 * threads / 32 * iters * ILP));
	}
	printf("\n");
}

int main() {
	unsigned* d_out; cudaMalloc(&d_out, 148 * 4 * 256 * 4);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	printf("SM cycles per warp instruction; m lanes share an address\n");
	run<0>("add 1, return (ATOMS.POPC.INC)", d_out, khz / 1000.0);
	run<1>("add 1, no return (RED)", d_out, khz / 1000.0);
	run<2>("packed add, return (ATOMS.ADD)", d_out, khz / 1000.0);
	run<3>("packed add, no return", d_out, khz / 1000.0);
	run<4>("run heads only add m (aggregated)", d_out, khz / 1000.0);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
