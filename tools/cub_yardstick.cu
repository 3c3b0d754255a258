// tools/cub_yardstick.cu -- YARDSTICK ONLY, never linked into libcl_ops.so: what NVIDIA's own CUB
// (the toolkit's headers) reaches on this B200 for the three BASELINE.json sort/scan shapes, so that
// the library's numbers can be read against a known-good implementation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cub_yardstick tools/cub_yardstick.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void fill(uint32_t* p, size_t n, uint32_t seed, uint32_t mask) {
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
		uint64_t x = i + seed * 0x9E3779B97F4A7C15ull;
		x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
		p[i] = (uint32_t) x & mask;
	}
}

template <typename F> static float best_ms(F f, int reps = 7) {
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	float best = 1e9f;
	for (int r = 0; r < reps; ++r) {
		cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
		float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 1 && ms < best) best = ms;
	}
	return best;
}

int main(int argc, char** argv) {
	const double peak = argc > 1 ? atof(argv[1]) : 6542.1;     // GB/s, MEASURED_PEAKS.json
	{   // C2: 2^28 u32 keys
		const size_t n = (size_t) 1 << 28;
		uint32_t *a, *b, *src; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&src, n * 4));
		fill<<<1184, 256>>>(src, n, 1, 0xffffffffu);
		size_t tb = 0; cub::DeviceRadixSort::SortKeys(nullptr, tb, a, b, n);
		void* tmp; CK(cudaMalloc(&tmp, tb));
		float ms = best_ms([&] { cudaMemcpyAsync(a, src, n * 4, cudaMemcpyDeviceToDevice); });
		float ms2 = best_ms([&] { cudaMemcpyAsync(a, src, n * 4, cudaMemcpyDeviceToDevice); cub::DeviceRadixSort::SortKeys(tmp, tb, a, b, n); });
		printf("{\"yardstick\": \"cub::DeviceRadixSort::SortKeys\", \"n\": %zu, \"dtype\": \"u32\", \"ms\": %.4f, \"gkeys_s\": %.2f, \"frac_of_36B_per_key\": %.3f, \"copy_gbs\": %.1f}\n",
			n, ms2 - ms, n / (ms2 - ms) * 1e-6, 36.0 * n / (ms2 - ms) * 1e-6 / peak, 8.0 * n / ms * 1e-6);
		cudaFree(a); cudaFree(b); cudaFree(src); cudaFree(tmp);
	}
	{   // C3 shape: 2^27 (u64 key, u32 payload)
		const size_t n = (size_t) 1 << 27;
		uint64_t *ka, *kb, *src; uint32_t *va, *vb;
		CK(cudaMalloc(&ka, n * 8)); CK(cudaMalloc(&kb, n * 8)); CK(cudaMalloc(&src, n * 8)); CK(cudaMalloc(&va, n * 4)); CK(cudaMalloc(&vb, n * 4));
		fill<<<1184, 256>>>((uint32_t*) src, n * 2, 2, 0xffffffffu);
		size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, ka, kb, va, vb, n);
		void* tmp; CK(cudaMalloc(&tmp, tb));
		float ms = best_ms([&] { cudaMemcpyAsync(ka, src, n * 8, cudaMemcpyDeviceToDevice); });
		float ms2 = best_ms([&] { cudaMemcpyAsync(ka, src, n * 8, cudaMemcpyDeviceToDevice); cub::DeviceRadixSort::SortPairs(tmp, tb, ka, kb, va, vb, n); });
		printf("{\"yardstick\": \"cub::DeviceRadixSort::SortPairs\", \"n\": %zu, \"dtype\": \"u64+u32\", \"ms\": %.4f, \"gpairs_s\": %.2f, \"frac_of_200B_per_pair\": %.3f}\n",
			n, ms2 - ms, n / (ms2 - ms) * 1e-6, 200.0 * n / (ms2 - ms) * 1e-6 / peak);
		cudaFree(ka); cudaFree(kb); cudaFree(src); cudaFree(va); cudaFree(vb); cudaFree(tmp);
	}
	{   // C4: exclusive scan of 2^30 u32 and f32
		const size_t n = (size_t) 1 << 30;
		uint32_t *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
		fill<<<1184, 256>>>(a, n, 3, 127u);
		size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, a, b, n);
		void* tmp; CK(cudaMalloc(&tmp, tb));
		float ms = best_ms([&] { cub::DeviceScan::ExclusiveSum(tmp, tb, a, b, n); });
		printf("{\"yardstick\": \"cub::DeviceScan::ExclusiveSum\", \"n\": %zu, \"dtype\": \"u32\", \"ms\": %.4f, \"gbs\": %.1f, \"frac\": %.3f}\n", n, ms, 8.0 * n / ms * 1e-6, 8.0 * n / ms * 1e-6 / peak);
		float msf = best_ms([&] { cub::DeviceScan::ExclusiveSum(tmp, tb, (float*) a, (float*) b, n); });
		printf("{\"yardstick\": \"cub::DeviceScan::ExclusiveSum\", \"n\": %zu, \"dtype\": \"f32\", \"ms\": %.4f, \"gbs\": %.1f, \"frac\": %.3f}\n", n, msf, 8.0 * n / msf * 1e-6, 8.0 * n / msf * 1e-6 / peak);
	}
	return 0;
}
