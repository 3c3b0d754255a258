// tools/ubench_atoms_zipf.cu -- cost of the placement atomic (packed add with return, warp-private
// rows) when the digits follow the Zipf(1.0)-over-2^20-hashed-values key distribution of
// tools/skew_bench.py, against uniform digits.  SM cycles per warp instruction, 32 warps per SM.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, const unsigned char* __restrict__ T, int iters) {
	__shared__ __align__(16) unsigned tab[8 * 256];
	__shared__ unsigned char sT[16384];
	for (int i = threadIdx.x; i < 16384; i += 256) sT[i] = T[i];
	const int warp = threadIdx.x >> 5;
	for (int i = threadIdx.x; i < 8 * 256; i += 256) tab[i] = 0;
	__syncthreads();
	unsigned* wt = tab + warp * 256;
	unsigned x[8];
	for (int i = 0; i < 8; ++i) x[i] = (1u + threadIdx.x * 7919u + blockIdx.x * 104729u + i * 31u) * 2654435761u;
	unsigned acc = 0;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			const unsigned d = sT[x[i] >> 18];
			if (MODE == 0) acc += atomicAdd(&wt[d >> 1], (d & 1) ? 65536u : 1u);
			else if (MODE == 1) atomicAdd(&wt[d >> 1], (d & 1) ? 65536u : 1u);
			else if (MODE == 2) atomicAdd(&wt[d], 1u);
			else acc += d;
			x[i] = x[i] * 1664525u + 1013904223u;
		}
	}
	out[threadIdx.x + blockIdx.x * blockDim.x] = acc + tab[threadIdx.x];
}
int main() {
	const int TN = 1 << 20;
	std::vector<unsigned char> zipf(TN), unif(TN);
	srand(1);
	for (int i = 0; i < TN; ++i) {
		double u = (rand() + 0.5) / (RAND_MAX + 1.0);
		unsigned long long rank = (unsigned long long) exp(u * log((double) (1 << 20)));
		if (rank < 1) rank = 1;
		unsigned key = (unsigned) ((rank * 2654435761ull) & 0xffffffffu);
		zipf[i] = (unsigned char) (key >> 8);      // digit of pass 1
		unif[i] = (unsigned char) (rand() >> 7);
	}
	unsigned char *dz, *du; unsigned* d_out;
	cudaMalloc(&dz, TN); cudaMalloc(&du, TN); cudaMalloc(&d_out, 148 * 4 * 256 * 4);
	cudaMemcpy(dz, zipf.data(), TN, cudaMemcpyHostToDevice); cudaMemcpy(du, unif.data(), TN, cudaMemcpyHostToDevice);
	int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	auto run = [&](auto kern, const char* name, const unsigned char* T) {
		const int iters = 2000, blocks = 148 * 4, threads = 256;
		cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
		kern<<<blocks, threads>>>(d_out, T, 10);
		cudaEventRecord(a); kern<<<blocks, threads>>>(d_out, T, iters); cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b);
		printf("%-44s %6.2f\n", name, ms * 1e-3 * (khz / 1000.0) * 1e6 / (4.0 * threads / 32 * iters * 8));
	};
	run(k<3>, "table lookup only (baseline), uniform", du);
	run(k<0>, "packed add return, uniform digits", du);
	run(k<0>, "packed add return, zipf digits", dz);
	run(k<1>, "packed add no return, uniform digits", du);
	run(k<1>, "packed add no return, zipf digits", dz);
	run(k<2>, "add 1 no return (POPC.INC), uniform", du);
	run(k<2>, "add 1 no return (POPC.INC), zipf", dz);
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
