/*
 * sort_common.h -- shared between the sorter object (sort.cu) and the sort
 * kernels (radix.cu, bitonic.cu): the key-extraction spec standing in for the
 * reference's CLO_SORT_KEY_GET / CLO_SORT_COMPARE macro strings
 * (/root/reference/src/cl_ops/sort/clo_sort_abstract.c:144-168).
 */
#ifndef CLO_SORT_COMMON_H
#define CLO_SORT_COMMON_H

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

enum { CLO_KIND_UNSIGNED = 0, CLO_KIND_SIGNED = 1, CLO_KIND_FLOAT = 2 };

/* key = (KEY_TYPE) (((x) >> shift) & mask), x of the element type */
struct CloKeySpec {
	uint64_t mask;        /* all ones: no mask */
	uint32_t shift;
	uint32_t elem_bits;   /* 8 * sizeof(elem) */
	uint32_t key_bits;    /* 8 * sizeof(key) */
	int elem_signed;      /* arithmetic >> and sign extension of the element */
	int key_kind;         /* CLO_KIND_* of the key type */
	int descending;       /* compare "((a) < (b))" */
	int identity;         /* shift == 0, no mask, key type == elem type */
};

#ifdef __CUDACC__
/* raw element bits (zero extended) -> raw key bits (zero extended, key_bits wide) */
__host__ __device__ inline uint64_t clo_extract_key(uint64_t raw, const CloKeySpec& ks) {
	if (ks.identity) return raw;
	uint64_t x = raw;
	if (ks.elem_signed && ks.elem_bits < 64) {
		const int sh = 64 - (int) ks.elem_bits;
		x = (uint64_t) (((int64_t) (raw << sh)) >> sh);
	}
	if (ks.elem_signed) x = (uint64_t) (((int64_t) x) >> ks.shift);
	else x = x >> ks.shift;
	x &= ks.mask;
	if (ks.key_bits < 64) x &= ((1ull << ks.key_bits) - 1);
	return x;
}

/* raw key bits -> unsigned 64-bit value whose natural order is the typed
 * "a > b" order of the key type (-0.0 == +0.0; NaN keys are not supported) */
__host__ __device__ inline uint64_t clo_ordered_key(uint64_t k, const CloKeySpec& ks) {
	if (ks.key_kind == CLO_KIND_SIGNED) {
		k ^= (1ull << (ks.key_bits - 1));
	} else if (ks.key_kind == CLO_KIND_FLOAT) {
		const uint64_t sign = 1ull << (ks.key_bits - 1);
		const uint64_t all = ks.key_bits < 64 ? ((1ull << ks.key_bits) - 1) : ~0ull;
		if (k == sign) k = 0;                       /* -0.0 -> +0.0 */
		k = (k & sign) ? (~k & all) : (k | sign);
	}
	return ks.descending ? ~k : k;
}
#endif

/* per-sorter scratch for the radix path, defined in radix.cu */
struct CloRadixState;
CloRadixState* clo_radix_state_new();
void clo_radix_state_free(CloRadixState* st);

/* Sort `n` elements of `elem_size` bytes (1,2,4,8) by the low `sorted_bits` bits
 * of the promoted key (see radix.cu).  src is never written unless src == dst
 * (in place).  payload_* may be NULL (keys only); payload is 32-bit.
 * Returns cudaSuccess or an error; *err_msg names unsupported configurations. */
cudaError_t clo_radix_sort(CloRadixState* st, int sm_count, size_t elem_size, const CloKeySpec& ks,
	uint32_t sorted_bits, const void* src, void* dst, const uint32_t* payload_src, uint32_t* payload_dst,
	size_t n, cudaStream_t stream, const char** err_msg);

/* Stable multi-way partition by splitters (sample sort); see radix.cu. */
cudaError_t clo_radix_partition(CloRadixState* st, size_t elem_size, const void* keys_in,
	const uint32_t* payload_in, void* keys_out, uint32_t* payload_out, size_t n, uint64_t gidx0,
	const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out,
	cudaStream_t stream, const char** err_msg);

/* status words of the last radix call: [0] look-back timeout flag, [1] tiles repaired by
 * the ballot path, [2..17] phase-profile cycle counters; synchronises the stream */
int clo_radix_debug(CloRadixState* st, cudaStream_t stream, unsigned long long out[18]);

/* per-kernel CUDA-event timing of the radix calls (bench evidence) */
void clo_radix_set_timing(CloRadixState* st, int on);
int clo_radix_get_timing(CloRadixState* st, float* out_ms, int cap);

/* device status flag of the last radix call (0 = ok); synchronises the stream */
int clo_radix_status(CloRadixState* st, cudaStream_t stream);

/* One onesweep pass (keys, or keys + u32 payload) with the persistent v6 kernel (radix_v6.cu).
 * elem_size 4 or 8; wide selects 64-bit AGG/PREF words; tile must equal the kernel's tile. */
cudaError_t clo_radix_v6_pass(int elem_size, int wide, int tile, const void* in, void* out,
		const uint32_t* payload_in, uint32_t* payload_out, size_t n,
		void* agg, void* pref, uint32_t* ticket, const unsigned long long* bins, uint32_t start_bit,
		uint32_t dmask, int* err, int sm_count, int prof_on, int flags, cudaStream_t stream,
		const void* chain_in = nullptr, void* chain_out = nullptr, void* out_alt = nullptr, uint32_t* vout_alt = nullptr);
/* the chain words (16 bytes each) say where the keys are between passes: a pass whose digit is
 * the same for every key moves nothing; the fixup copies the result to dst only when needed */
cudaError_t clo_radix_v6_chain_fixup(const void* chain, void* dst, void* vdst, size_t key_bytes, size_t val_bytes,
		int sm_count, cudaStream_t stream);

/* bitonic / gselect, defined in bitonic.cu */
struct CloBitonicState;
CloBitonicState* clo_bitonic_state_new();
void clo_bitonic_state_free(CloBitonicState* st);
cudaError_t clo_bitonic_sort(CloBitonicState* st, size_t elem_size, const CloKeySpec& ks,
	void* data, size_t n, cudaStream_t stream);
/* abitonic's tuning options: steps fused in registers (maxps, 1..4) and the largest number of
 * steps taken in shared memory (maxsfs; 13 = the full 8192-element tile) */
void clo_bitonic_set_fusion(CloBitonicState* st, int max_private_steps, int max_local_steps);
cudaError_t clo_gselect_sort(size_t elem_size, const CloKeySpec& ks, const void* in, void* out,
	size_t n, cudaStream_t stream);

/* the two stages of the partition, for the fused partition + exchange of the sample sort:
 * count -> (the ranks exchange their bucket sizes) -> scatter straight into the receive
 * buffers of the destination ranks (dests[q] may be peer memory) */
cudaError_t clo_radix_partition_count(CloRadixState* st, size_t elem_size, const void* keys_in, size_t n, uint64_t gidx0,
		const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts, uint64_t* counts_out,
		cudaStream_t stream, const char** err_msg);
cudaError_t clo_radix_partition_scatter(CloRadixState* st, size_t elem_size, const void* keys_in, const uint32_t* payload_in,
		size_t n, uint64_t gidx0, const void* splitter_keys, const uint64_t* splitter_idx, uint32_t nparts,
		const uint64_t* first_slot, void* const* dests, void* const* vdests, const int* ok,
		cudaStream_t stream, const char** err_msg);

/* run-time compiled comparison sorts for macro strings outside the precompiled menu (jit.cu) */
#ifdef __cplusplus
#include <string>
struct CloJitSort;
CloJitSort* clo_jit_sort_new(CloType elem_type, CloType key_type, const char* compare, const char* get_key, std::string& msg);
void clo_jit_sort_free(CloJitSort* j);
const char* clo_jit_sort_source(CloJitSort* j);
cudaError_t clo_jit_bitonic_sort(CloJitSort* j, size_t elem_size, void* data, size_t n, cudaStream_t stream);
cudaError_t clo_jit_gselect_sort(CloJitSort* j, const void* in, void* out, size_t n, cudaStream_t stream);
/* satradix with a run-time compiled get_key: (key, index) extraction, pair sort, gather */
cudaError_t clo_jit_radix_sort(CloJitSort* j, CloRadixState* rs, int sm_count, size_t elem_size, size_t key_size,
	uint32_t sorted_bits, const void* in, void* out, size_t n, cudaStream_t stream, const char** err_msg);
#endif

#endif
