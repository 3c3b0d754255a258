/*
 * sort.cu -- the sorter object and its four implementations behind the
 * reference's API (/root/reference/src/cl_ops/sort/clo_sort_abstract.c:91-629,
 * clo_sort_{sbitonic,abitonic,gselect,satradix}.c).
 */
#include "clo_internal.h"
#include "sort_common.h"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <string>

struct clo_sort {
	CloSortImplDef impl_def;
	CCLContext* ctx;
	CCLProgram* prg;
	CloType elem_type;
	CloType key_type;
	void* data;
	/* backend */
	CloKeySpec ks;
	unsigned radix;          /* satradix "radix=" option (clo_sort_satradix.c:352,385-392) */
	int typed_order;         /* satradix "typed_order=1": opt-in numeric order for signed / float keys */
	unsigned minps, maxps, maxsfs;  /* abitonic options: maxps / maxsfs set the fusion depth of the bitonic kernel */
	CloRadixState* rs;
	CloBitonicState* bs;
	CloJitSort* jit;         /* run-time compiled network for compare / get_key strings outside the menu */
	CloScratch host_in, host_out;   /* device staging of clo_sort_with_host_data, kept between calls */
};

static ccl_program g_sort_program = { "clo_sort (precompiled sm_100a)", nullptr, std::string(), nullptr, {} };

/* ------------------------------------------------ macro-string "compilation" */

static std::string squeeze(const char* s) {
	std::string o;
	for (; s && *s; ++s) if (!isspace((unsigned char) *s) && *s != '(' && *s != ')') o.push_back(*s);
	return o;
}

static bool parse_uint(const std::string& s, size_t& i, uint64_t& v) {
	size_t start = i;
	int base = 10;
	if (s.compare(i, 2, "0x") == 0 || s.compare(i, 2, "0X") == 0) { base = 16; i += 2; start = i; }
	v = 0;
	while (i < s.size() && isxdigit((unsigned char) s[i]) && (base == 16 || isdigit((unsigned char) s[i]))) {
		int d = isdigit((unsigned char) s[i]) ? s[i] - '0' : (tolower(s[i]) - 'a' + 10);
		v = v * base + (uint64_t) d;
		++i;
	}
	if (i == start) return false;
	while (i < s.size() && (s[i] == 'u' || s[i] == 'U' || s[i] == 'l' || s[i] == 'L')) ++i;
	return true;
}

/* get_key menu: x | x>>K | x&M | x>>K&M  (after removing blanks and parentheses) */
static bool parse_get_key(const char* get_key, uint32_t& shift, uint64_t& mask) {
	shift = 0; mask = ~0ull;
	if (!get_key) return true;
	std::string s = squeeze(get_key);
	size_t i = 0;
	if (s.empty() || s[i] != 'x') return false;
	++i;
	if (s.compare(i, 2, ">>") == 0) {
		i += 2;
		uint64_t k;
		if (!parse_uint(s, i, k) || k > 63) return false;
		shift = (uint32_t) k;
	}
	if (i < s.size() && s[i] == '&') {
		++i;
		if (!parse_uint(s, i, mask)) return false;
	}
	return i == s.size();
}

/* compare menu: a>b (ascending, the default) | a<b (descending) */
static bool parse_compare(const char* compare, int& descending) {
	descending = 0;
	if (!compare) return true;
	std::string s = squeeze(compare);
	if (s == "a>b") return true;
	if (s == "a<b") { descending = 1; return true; }
	return false;
}

static int kind_of(CloType t) {
	switch (t) {
	case CLO_CHAR: case CLO_SHORT: case CLO_INT: case CLO_LONG: return CLO_KIND_SIGNED;
	case CLO_FLOAT: case CLO_DOUBLE: return CLO_KIND_FLOAT;
	default: return CLO_KIND_UNSIGNED;
	}
}

/* ------------------------------------------------------------ option strings */

/* split "k=v,k=v"; returns false on a token that is not key=value */
static bool next_option(const char*& p, std::string& key, std::string& val, bool& bad) {
	bad = false;
	while (*p == ',') ++p;
	if (!*p) return false;
	const char* e = strchr(p, ',');
	std::string tok = e ? std::string(p, e - p) : std::string(p);
	p = e ? e : p + strlen(p);
	size_t eq = tok.find('=');
	if (eq == std::string::npos) { bad = true; key = tok; return true; }
	key = tok.substr(0, eq); val = tok.substr(eq + 1);
	return true;
}

/* ------------------------------------------------------------- satradix */

static const char* satradix_init(CloSort* sorter, const char* options, GError** err) {
	sorter->radix = 16;
	sorter->typed_order = 0;
	if (options) {
		const char* p = options;
		std::string k, v; bool bad;
		while (next_option(p, k, v, bad)) {
			if (bad) {
				g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Invalid option '%s' for abitonic sort.", k.c_str());
				return NULL;
			}
			if (k == "radix") {
				sorter->radix = (unsigned) atoi(v.c_str());
				if (clo_ones32(sorter->radix) != 1) {
					g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Radix must be a power of 2.");
					return NULL;
				}
			} else if (k == "typed_order") {
				/* not in the reference, which sorts raw key bits whatever the key type
				 * (clo_sort_satradix.cl:61: negative ints after positive ones, floats by bit
				 * pattern).  Opt-in: order signed integers and floats by VALUE. */
				sorter->typed_order = atoi(v.c_str()) != 0;
			} else if (k.size() >= 4 && strncasecmp(k.c_str(), "scan", 4) == 0) {
				/* the reference forwards these to its internal scanner
				 * (clo_sort_satradix.c:393-406); the onesweep has no scan kernel.
				 * The scan type is still validated. */
				if (k.size() == 4 && v != "blelloch") {
					g_set_error(err, CLO_ERROR, CLO_ERROR_IMPL_NOT_FOUND,
						"The requested scan implementation, '%s', was not found.", v.c_str());
					return NULL;
				}
			} else {
				g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Invalid option key '%s' for satradix sort.", k.c_str());
				return NULL;
			}
		}
	}
	return "";
}

static void generic_finalize(CloSort* sorter) { (void) sorter; }

static bool check_buffers(CloSort* sorter, CCLBuffer* in, CCLBuffer* out, size_t numel, GError** err) {
	const size_t bytes = numel * clo_type_sizeof(sorter->elem_type);
	if (!in || in->size < bytes || (out && out->size < bytes)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "sort: buffers too small for %zu elements", numel);
		return false;
	}
	return true;
}

/* Bits of the promoted key the reference's passes actually sort on:
 * total_digits * bits_in_digit with total_digits = elem_bits / bits_in_digit
 * (clo_sort_satradix.c:166-169); shift counts wrap at the promoted key width. */
static bool satradix_sorted_bits(CloSort* sorter, uint32_t& bits, GError** err) {
	const unsigned nb = clo_tzc((int) sorter->radix);
	if (nb == 0) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Radix must be at least 2."); return false; }
	const unsigned elem_bits = 8 * (unsigned) clo_type_sizeof(sorter->elem_type);
	const unsigned W = clo_type_sizeof(sorter->key_type) == 8 ? 64 : 32;
	unsigned sorted = (elem_bits / nb) * nb;
	if (sorted > W) {
		if (W % nb) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS,
				"satradix: radix %u with a %u-bit element and a %u-bit key is not supported", sorter->radix, elem_bits, W);
			return false;
		}
		sorted = W;
	}
	if (sorted == 0) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "satradix: radix wider than the element"); return false; }
	bits = sorted;
	return true;
}

/* typed_order: a bijection on the key bits that turns numeric order into unsigned order (sign
 * bit flipped for signed integers; floats: all bits of negatives, the sign bit of the others)
 * and its inverse, applied around the raw-bit passes */
template <typename U, bool IS_FLOAT, bool INVERSE>
__global__ void clo_radix_typed_flip(const U* __restrict__ in, U* __restrict__ out, size_t n) {
	const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const U sign = (U) 1 << (8 * sizeof(U) - 1);
	U k = in[i];
	if (!IS_FLOAT) k ^= sign;
	else if (!INVERSE) k ^= (k & sign) ? (U) ~(U) 0 : sign;
	else k ^= (k & sign) ? sign : (U) ~(U) 0;
	out[i] = k;
}

template <typename U>
static void typed_flip(bool is_float, bool inverse, const void* in, void* out, size_t n, cudaStream_t stream) {
	const unsigned grid = (unsigned) ((n + 255) / 256);
	if (!is_float) clo_radix_typed_flip<U, false, false><<<grid, 256, 0, stream>>>((const U*) in, (U*) out, n);
	else if (!inverse) clo_radix_typed_flip<U, true, false><<<grid, 256, 0, stream>>>((const U*) in, (U*) out, n);
	else clo_radix_typed_flip<U, true, true><<<grid, 256, 0, stream>>>((const U*) in, (U*) out, n);
	CLO_COUNT_LAUNCH(1);
}

static void typed_flip_any(size_t es, bool is_float, bool inverse, const void* in, void* out, size_t n, cudaStream_t stream) {
	switch (es) {
	case 1: typed_flip<unsigned char>(false, inverse, in, out, n, stream); break;
	case 2: typed_flip<unsigned short>(false, inverse, in, out, n, stream); break;
	case 4: typed_flip<unsigned int>(is_float, inverse, in, out, n, stream); break;
	default: typed_flip<unsigned long long>(is_float, inverse, in, out, n, stream); break;
	}
}

static CCLEvent* satradix_sort_with_device_data(CloSort* sorter, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	(void) cq_comm; (void) lws_max;
	if ((err && *err) || !cq_exec) return NULL;
	if (!check_buffers(sorter, data_in, data_out, numel, err)) return NULL;
	const bool typed = sorter->typed_order && sorter->ks.key_kind != CLO_KIND_UNSIGNED;
	if (typed && (sorter->jit || !sorter->ks.identity || sorter->key_type != sorter->elem_type)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "satradix typed_order needs the identity key (key type = element type, get_key (x))");
		return NULL;
	}
	if (sorter->ks.key_kind == CLO_KIND_FLOAT && !typed) {
		/* `key >> b` does not compile for a float key in OpenCL C
		 * (clo_sort_satradix.cl:61): the reference fails at build time */
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "satradix sorts integer keys only");
		return NULL;
	}
	uint32_t bits = 0;
	if (!satradix_sorted_bits(sorter, bits, err)) return NULL;
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_radix_onesweep");
	const char* msg = NULL;
	void* dst = data_out ? data_out->ptr : data_in->ptr;
	if (typed) {
		/* flip -> raw-bit passes in place -> flip back */
		const size_t es = clo_type_sizeof(sorter->elem_type);
		const bool is_float = sorter->ks.key_kind == CLO_KIND_FLOAT;
		typed_flip_any(es, is_float, false, data_in->ptr, dst, numel, cq_exec->stream);
		cudaError_t rc2 = clo_radix_sort(sorter->rs, clo_sm_count(cq_exec->ctx->dev.ordinal), es, sorter->ks,
			(uint32_t) (8 * es), dst, dst, NULL, NULL, numel, cq_exec->stream, &msg);
		if (rc2 == cudaSuccess) { typed_flip_any(es, is_float, true, dst, dst, numel, cq_exec->stream); rc2 = cudaGetLastError(); }
		clo_queue_end(cq_exec, evt);
		if (rc2 != cudaSuccess && msg) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg); return NULL; }
		if (clo_cuda_failed(rc2, err, "clo_radix_sort (typed order)")) return NULL;
		return evt;
	}
	cudaError_t rc = sorter->jit
		? clo_jit_radix_sort(sorter->jit, sorter->rs, clo_sm_count(cq_exec->ctx->dev.ordinal),
			clo_type_sizeof(sorter->elem_type), clo_type_sizeof(sorter->key_type), bits, data_in->ptr, dst, numel, cq_exec->stream, &msg)
		: clo_radix_sort(sorter->rs, clo_sm_count(cq_exec->ctx->dev.ordinal),
			clo_type_sizeof(sorter->elem_type), sorter->ks, bits, data_in->ptr, dst, NULL, NULL,
			numel, cq_exec->stream, &msg);
	clo_queue_end(cq_exec, evt);
	if (rc != cudaSuccess && msg) {
		g_set_error(err, CLO_ERROR, rc == cudaErrorLaunchFailure ? CLO_ERROR_LIBRARY : CLO_ERROR_ARGS, "%s", msg);
		return NULL;
	}
	if (clo_cuda_failed(rc, err, "clo_radix_sort")) return NULL;
	return evt;
}

static const char* const kSatradixKernels[] = { "clo_radix_histogram", "clo_radix_scan_bins", "clo_radix_onesweep_v6" };

static cl_uint satradix_get_num_kernels(CloSort* s, GError** err) { (void) s; (void) err; return 3; }

static const char* satradix_get_kernel_name(CloSort* s, cl_uint i, GError** err) {
	(void) s;
	if (i >= 3) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return NULL; }
	return kSatradixKernels[i];
}

static size_t satradix_get_localmem_usage(CloSort* s, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	(void) lws_max; (void) numel;
	/* shared memory of the kernels as launched for a keys-only sort of this element type:
	 * histogram [passes][256 bins][columns] u32 (radix.cu launch_histogram); the bin scan's warp
	 * totals; the onesweep pass' packed warp rows, digit-start / offset tables and its two staging
	 * buffers of one tile each (radix_v6.cuh onesweep_v6_smem; 512 x 16 keys of 4 bytes, 512 x 10
	 * of 8 bytes; narrower keys run the older pass with one staging buffer of 512 x 16 keys) */
	const size_t es = clo_type_sizeof(s->elem_type);
	const size_t passes = es;                          /* 8-bit digits */
	switch (i) {
	case 0: return (passes <= 2 ? passes : (passes <= 4 ? 4 : 8)) * 256 * (passes <= 4 ? 32 : 16) * 4;
	case 1: return 8 * 8;
	case 2:
		if (es >= 4) return 16 * 128 * 4 + 4 * 256 * 4 + 2 * 256 * 8 + 32 * 4 + 2 * (size_t) 512 * (es == 8 ? 10 : 16) * es;
		return 16 * 256 * 4 + 256 * 4 + 256 * 8 + 16 * 4 + (size_t) 512 * 16 * es;
	default: g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return 0;
	}
}

extern "C" const CloSortImplDef clo_sort_satradix_def = {
	"satradix", CL_TRUE, satradix_init, generic_finalize, satradix_sort_with_device_data,
	satradix_get_num_kernels, satradix_get_kernel_name, satradix_get_localmem_usage
};

/* -------------------------------------------------------- s/a-bitonic */

static const char* sbitonic_init(CloSort* sorter, const char* options, GError** err) {
	/* clo_sort_sbitonic.c:140-150 ignores its options */
	(void) sorter; (void) options; (void) err;
	return "";
}

static const char* abitonic_init(CloSort* sorter, const char* options, GError** err) {
	/* clo_sort_abitonic.c:459-542: minps/maxps in [1,4], maxsfs free, minps <= maxps */
	sorter->maxps = 4; sorter->minps = 1; sorter->maxsfs = 0xffffffffu;
	if (options) {
		const char* p = options;
		std::string k, v; bool bad;
		while (next_option(p, k, v, bad)) {
			if (bad) {
				g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Invalid option '%s' for abitonic sort.", k.c_str());
				return NULL;
			}
			const unsigned value = (unsigned) atoi(v.c_str());
			if (k == "minps") {
				if (value > 4 || value < 1) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Option 'minps' must be between 1 and 4."); return NULL; }
				sorter->minps = value;
			} else if (k == "maxps") {
				if (value > 4 || value < 1) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Option 'maxps' must be between 1 and 4."); return NULL; }
				sorter->maxps = value;
			} else if (k == "maxsfs") {
				sorter->maxsfs = value;
			} else {
				g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "Invalid option key '%s' for abitonic sort.", k.c_str());
				return NULL;
			}
		}
		if (sorter->maxps < sorter->minps) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "'minps' (%d) must be less or equal than 'maxps' (%d).",
				(int) sorter->minps, (int) sorter->maxps);
			return NULL;
		}
	}
	return "";
}

static CCLEvent* bitonic_sort_with_device_data(CloSort* sorter, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	(void) cq_comm; (void) lws_max;
	if ((err && *err) || !cq_exec) return NULL;
	if (!check_buffers(sorter, data_in, data_out, numel, err)) return NULL;
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	const size_t bytes = numel * clo_type_sizeof(sorter->elem_type);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_bitonic");
	cudaError_t rc = cudaSuccess;
	void* work = data_in->ptr;
	if (data_out && data_out->ptr != data_in->ptr) {
		/* clo_sort_sbitonic.c:83-96: copy, then sort the copy */
		rc = cudaMemcpyAsync(data_out->ptr, data_in->ptr, bytes, cudaMemcpyDeviceToDevice, cq_exec->stream);
		work = data_out->ptr;
	}
	if (rc == cudaSuccess && sorter->jit)
		rc = clo_jit_bitonic_sort(sorter->jit, clo_type_sizeof(sorter->elem_type), work, numel, cq_exec->stream);
	else if (rc == cudaSuccess)
		rc = clo_bitonic_sort(sorter->bs, clo_type_sizeof(sorter->elem_type), sorter->ks, work, numel, cq_exec->stream);
	clo_queue_end(cq_exec, evt);
	if (clo_cuda_failed(rc, err, "clo_bitonic_sort")) return NULL;
	return evt;
}

static const char* const kBitonicKernels[] = { "clo_bitonic_fused", "clo_bitonic_local", "clo_bitonic_global" };

static cl_uint bitonic_get_num_kernels(CloSort* s, GError** err) { (void) s; (void) err; return 3; }

static const char* bitonic_get_kernel_name(CloSort* s, cl_uint i, GError** err) {
	(void) s;
	if (i >= 3) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return NULL; }
	return kBitonicKernels[i];
}

/* shared memory of the kernels above: the fused kernel's padded tile (8192 elements of up to
 * 4 bytes, 4096 of 8 bytes, one padding element per 32); the padded-N kernels' 4096-element tile
 * plus its flag bytes; the global step uses none */
static size_t bitonic_get_localmem_usage(CloSort* s, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	(void) lws_max; (void) numel;
	if (i >= 3) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return 0; }
	const size_t es = clo_type_sizeof(s->elem_type);
	const size_t tile = es == 8 ? 4096 : 8192;
	if (i == 0) return (tile + tile / 32 + 1) * es;
	return i == 1 ? 4096 * es + 4096 : 0;
}

extern "C" const CloSortImplDef clo_sort_sbitonic_def = {
	"sbitonic", CL_TRUE, sbitonic_init, generic_finalize, bitonic_sort_with_device_data,
	bitonic_get_num_kernels, bitonic_get_kernel_name, bitonic_get_localmem_usage
};

extern "C" const CloSortImplDef clo_sort_abitonic_def = {
	"abitonic", CL_TRUE, abitonic_init, generic_finalize, bitonic_sort_with_device_data,
	bitonic_get_num_kernels, bitonic_get_kernel_name, bitonic_get_localmem_usage
};

/* -------------------------------------------------------------- gselect */

static const char* gselect_init(CloSort* sorter, const char* options, GError** err) {
	/* clo_sort_gselect.c:150-160 ignores its options */
	(void) sorter; (void) options; (void) err;
	return "";
}

static CCLEvent* gselect_sort_with_device_data(CloSort* sorter, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	(void) cq_comm; (void) lws_max;
	if ((err && *err) || !cq_exec) return NULL;
	if (!check_buffers(sorter, data_in, data_out, numel, err)) return NULL;
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	const size_t es = clo_type_sizeof(sorter->elem_type), bytes = numel * es;
	void* tmp = NULL;
	void* out = data_out ? data_out->ptr : NULL;
	if (!out || out == data_in->ptr) {
		/* clo_sort_gselect.c:60-70,117-124: no output buffer -> temporary + copy back */
		if (clo_cuda_failed(cudaMallocAsync(&tmp, bytes ? bytes : 1, cq_exec->stream), err, "cudaMallocAsync")) return NULL;
		out = tmp;
	}
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_gselect");
	cudaError_t rc = sorter->jit ? clo_jit_gselect_sort(sorter->jit, data_in->ptr, out, numel, cq_exec->stream)
		: clo_gselect_sort(es, sorter->ks, data_in->ptr, out, numel, cq_exec->stream);
	if (rc == cudaSuccess && tmp)
		rc = cudaMemcpyAsync(data_in->ptr, tmp, bytes, cudaMemcpyDeviceToDevice, cq_exec->stream);
	if (tmp) cudaFreeAsync(tmp, cq_exec->stream);
	clo_queue_end(cq_exec, evt);
	if (clo_cuda_failed(rc, err, "clo_gselect_sort")) return NULL;
	return evt;
}

static const char* const kGselectKernels[] = { "clo_gselect_kernel" };

static cl_uint gselect_get_num_kernels(CloSort* s, GError** err) { (void) s; (void) err; return 1; }

static const char* gselect_get_kernel_name(CloSort* s, cl_uint i, GError** err) {
	(void) s;
	if (i >= 1) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return NULL; }
	return kGselectKernels[i];
}

static size_t gselect_get_localmem_usage(CloSort* s, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	(void) s; (void) lws_max; (void) numel;
	if (i >= 1) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "kernel index %u out of range", i); return 0; }
	return 256 * 8;
}

extern "C" const CloSortImplDef clo_sort_gselect_def = {
	"gselect", CL_FALSE, gselect_init, generic_finalize, gselect_sort_with_device_data,
	gselect_get_num_kernels, gselect_get_kernel_name, gselect_get_localmem_usage
};

/* --------------------------------------------------------------- object */

extern "C" CloSort* clo_sort_new(const char* type, const char* options, CCLContext* ctx,
		CloType* elem_type, CloType* key_type, const char* compare, const char* get_key,
		const char* compiler_opts, GError** err) {
	(void) compiler_opts; /* OpenCL build options: accepted and ignored */
	if (err && *err) return NULL;
	if (!ctx || !elem_type) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "NULL context or element type"); return NULL; }
	const CloSortImplDef* impls[] = { &clo_sort_sbitonic_def, &clo_sort_abitonic_def,
		&clo_sort_gselect_def, &clo_sort_satradix_def };
	const CloSortImplDef* def = NULL;
	for (const CloSortImplDef* d : impls) if (type && strcmp(type, d->name) == 0) def = d;
	if (!def) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_IMPL_NOT_FOUND,
			"The requested sort implementation, '%s', was not found.", type ? type : "(null)");
		return NULL;
	}
	const CloType et = *elem_type, kt = key_type ? *key_type : *elem_type;
	if (clo_type_sizeof(et) == 0 || clo_type_sizeof(kt) == 0 || et == CLO_HALF || kt == CLO_HALF) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_UNKNOWN_TYPE, "Unsupported sort types (elem=%d, key=%d)", (int) et, (int) kt);
		return NULL;
	}
	clo_sort* s = new clo_sort();
	s->impl_def = *def;
	s->ctx = ctx; ccl_context_ref(ctx);
	s->prg = &g_sort_program;
	s->elem_type = et; s->key_type = kt;
	s->data = NULL;
	s->radix = 16; s->minps = 1; s->maxps = 4; s->maxsfs = 0xffffffffu;
	s->rs = clo_radix_state_new();
	s->bs = clo_bitonic_state_new();
	s->jit = NULL;

	GError* ierr = NULL;
	uint32_t shift = 0; uint64_t mask = ~0ull; int desc = 0;
	s->impl_def.init(s, options, &ierr);
	/* abitonic's tuning options are real knobs here: steps fused in registers (maxps) and steps
	 * taken in the shared-memory tile (maxsfs) -- clo_sort_abitonic.c:486-542 */
	if (!ierr && def == &clo_sort_abitonic_def)
		clo_bitonic_set_fusion(s->bs, (int) s->maxps, s->maxsfs > 13u ? 13 : (int) s->maxsfs);
	const bool comparison_sort = def != &clo_sort_satradix_def;
	/* satradix never looks at `compare` (the reference defines the macro and its radix kernels
	 * do not use it): raw key bits ascending whatever it says */
	bool in_menu = parse_get_key(get_key, shift, mask) && (parse_compare(compare, desc) || !comparison_sort);
	if (!comparison_sort) desc = 0;
	/* an integer element with a float key type (or the reverse) is a VALUE conversion in the
	 * reference, `(float) (x)`: the precompiled kernels only reinterpret bits, so these go through
	 * the run-time compiler, which emits the real cast */
	if (in_menu && ((kind_of(et) == CLO_KIND_FLOAT) != (kind_of(kt) == CLO_KIND_FLOAT))) {
		in_menu = false; shift = 0; mask = ~0ull;
	}
	if (!ierr && !in_menu && comparison_sort) {
		/* the reference compiles ANY macro body into its kernels (clo_sort_abstract.c:144-168);
		 * strings outside the precompiled menu are compiled here too, with NVRTC */
		std::string msg;
		CloDeviceGuard g(ctx->dev.ordinal);
		s->jit = clo_jit_sort_new(et, kt, compare, get_key, msg);
		if (!s->jit) g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg.c_str());
		shift = 0; mask = ~0ull; desc = 0;
	}
	if (!ierr && !in_menu && !comparison_sort) {
		/* a get_key string outside the menu (clo_sort_abstract.c:157-168 accepts any macro body;
		 * the digit is cut from it at clo_sort_satradix.cl:61): the key extraction is compiled at
		 * run time, the passes are the library's own (jit.cu: clo_jit_radix_sort) */
		std::string msg;
		CloDeviceGuard g(ctx->dev.ordinal);
		s->jit = clo_jit_sort_new(et, kt, NULL, get_key, msg);
		if (!s->jit) g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg.c_str());
		shift = 0; mask = ~0ull; desc = 0;
	}
	if (!ierr && !s->jit && kind_of(et) == CLO_KIND_FLOAT && (shift != 0 || mask != ~0ull || kt != et))
		g_set_error(&ierr, CLO_ERROR, CLO_ERROR_ARGS, "a floating-point element only supports the identity key");
	if (ierr) { g_propagate_error(err, ierr); clo_sort_destroy(s); return NULL; }

	CloKeySpec& ks = s->ks;
	ks.mask = mask; ks.shift = shift;
	ks.elem_bits = 8 * (uint32_t) clo_type_sizeof(et);
	ks.key_bits = 8 * (uint32_t) clo_type_sizeof(kt);
	ks.elem_signed = kind_of(et) == CLO_KIND_SIGNED;
	ks.key_kind = kind_of(kt);
	ks.descending = desc;
	ks.identity = (shift == 0 && mask == ~0ull && ks.key_bits == ks.elem_bits &&
		kind_of(et) != CLO_KIND_FLOAT) ? 1 : 0;
	if (kind_of(et) == CLO_KIND_FLOAT) ks.identity = 1;  /* raw bits are the key bits */
	clo_handle_add(s);
	return s;
}

extern "C" void clo_sort_destroy(CloSort* sorter) {
	if (!sorter) return;
	clo_handle_remove(sorter);
	sorter->impl_def.finalize(sorter);
	{
		CloDeviceGuard g(sorter->ctx->dev.ordinal);
		clo_radix_state_free(sorter->rs);
		clo_bitonic_state_free(sorter->bs);
		clo_jit_sort_free(sorter->jit);
		sorter->host_in.release();
		sorter->host_out.release();
	}
	ccl_context_unref(sorter->ctx);
	delete sorter;
}

extern "C" CCLEvent* clo_sort_with_device_data(CloSort* sorter, CCLQueue* cq_exec, CCLQueue* cq_comm,
		CCLBuffer* data_in, CCLBuffer* data_out, size_t numel, size_t lws_max, GError** err) {
	if (!sorter || (err && *err) || !cq_exec) return NULL;
	return sorter->impl_def.sort_with_device_data(sorter, cq_exec, cq_comm, data_in, data_out, numel, lws_max, err);
}

/* Host data in, sorted host data out (the contract of clo_sort_abstract.c:296-418: blocks until
 * data_out is complete).  The device staging lives in the sorter and only grows -- no
 * cudaMalloc / cudaFree of gigabytes per call -- and when one queue does everything the copy in,
 * the sort and the copy out are simply stream ordered: the host blocks once, at the end. */
extern "C" cl_bool clo_sort_with_host_data(CloSort* sorter, CCLQueue* cq_exec, CCLQueue* cq_comm,
		void* data_in, void* data_out, size_t numel, size_t lws_max, GError** err) {
	if (!sorter || (err && *err)) return CL_FALSE;
	if (numel && (!data_in || !data_out)) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "sort: NULL host pointer"); return CL_FALSE; }
	const size_t bytes = numel * clo_type_sizeof(sorter->elem_type);
	CCLQueue* own_queue = NULL;
	if (!cq_exec) {
		own_queue = ccl_queue_new(sorter->ctx, NULL, 0, err);
		if (!own_queue) return CL_FALSE;
		cq_exec = own_queue;
	}
	if (!cq_comm) cq_comm = cq_exec;
	const bool one_queue = cq_comm == cq_exec;
	cl_bool ok = CL_FALSE;
	CCLBuffer *in_dev = NULL, *out_dev = NULL;
	CCLEventWaitList ewl = NULL;
	{
		CloDeviceGuard g(sorter->ctx->dev.ordinal);
		const bool two = !sorter->impl_def.in_place;
		if (clo_cuda_failed(sorter->host_in.reserve(bytes ? bytes : 1), err, "sort staging") ||
				(two && clo_cuda_failed(sorter->host_out.reserve(bytes ? bytes : 1), err, "sort staging"))) goto done;
		in_dev = ccl_buffer_new_wrap(sorter->ctx, sorter->host_in.ptr, bytes, err);
		if (in_dev && two) out_dev = ccl_buffer_new_wrap(sorter->ctx, sorter->host_out.ptr, bytes, err);
		if (!in_dev || (two && !out_dev)) goto done;
		CCLEvent* evt = ccl_buffer_enqueue_write(in_dev, cq_comm, CL_FALSE, 0, bytes, data_in, NULL, err);
		if (!evt) goto done;
		/* a separate copy queue must have delivered the data; one queue is already ordered */
		if (!one_queue && !ccl_queue_finish(cq_comm, err)) goto done;
		evt = sorter->impl_def.sort_with_device_data(sorter, cq_exec, cq_comm, in_dev, out_dev, numel, lws_max, err);
		if (!evt && numel) goto done;
		evt = ccl_buffer_enqueue_read(two ? out_dev : in_dev, cq_comm, CL_FALSE, 0, bytes, data_out,
			(!one_queue && evt) ? ccl_ewl(&ewl, evt, NULL) : NULL, err);
		if (!evt) goto done;
		if (!ccl_queue_finish(cq_comm, err)) goto done;                 /* the one blocking point */
		if (!one_queue && !ccl_queue_finish(cq_exec, err)) goto done;
		if (clo_radix_status(sorter->rs, cq_exec->stream) != 0) {
			g_set_error(err, CLO_ERROR, CLO_ERROR_LIBRARY, "radix look-back timed out (device status flag set)");
			goto done;
		}
		ok = CL_TRUE;
	}
done:
	ccl_event_wait_list_clear(&ewl);
	if (in_dev) ccl_buffer_destroy(in_dev);
	if (out_dev) ccl_buffer_destroy(out_dev);
	if (own_queue) ccl_queue_destroy(own_queue);
	return ok;
}

extern "C" CCLContext* clo_sort_get_context(CloSort* s) { return s ? s->ctx : NULL; }
extern "C" CCLProgram* clo_sort_get_program(CloSort* s) { return s ? s->prg : NULL; }
extern "C" CloType clo_sort_get_element_type(CloSort* s) { return s ? s->elem_type : (CloType) -1; }
extern "C" size_t clo_sort_get_element_size(CloSort* s) { return s ? clo_type_sizeof(s->elem_type) : 0; }
extern "C" CloType clo_sort_get_key_type(CloSort* s) { return s ? s->key_type : (CloType) -1; }
extern "C" size_t clo_sort_get_key_size(CloSort* s) { return s ? clo_type_sizeof(s->key_type) : 0; }
extern "C" void* clo_sort_get_data(CloSort* s) { return s ? s->data : NULL; }
extern "C" void clo_sort_set_data(CloSort* s, void* data) { if (s) s->data = data; }
extern "C" cl_uint clo_sort_get_num_kernels(CloSort* s, GError** err) { return s ? s->impl_def.get_num_kernels(s, err) : 0; }
extern "C" const char* clo_sort_get_kernel_name(CloSort* s, cl_uint i, GError** err) {
	return s ? s->impl_def.get_kernel_name(s, i, err) : NULL;
}
extern "C" size_t clo_sort_get_localmem_usage(CloSort* s, cl_uint i, size_t lws_max, size_t numel, GError** err) {
	return s ? s->impl_def.get_localmem_usage(s, i, lws_max, numel, err) : 0;
}

/* ------------------------------------------------ additive entry points */

extern "C" void clo_sort_b200_set_timing(CloSort* sorter, cl_bool on) {
	if (sorter) clo_radix_set_timing(sorter->rs, on ? 1 : 0);
}

extern "C" cl_uint clo_sort_b200_get_timing(CloSort* sorter, float* out_ms, cl_uint cap) {
	if (!sorter || !out_ms) return 0;
	CloDeviceGuard g(sorter->ctx->dev.ordinal);
	return (cl_uint) clo_radix_get_timing(sorter->rs, out_ms, (int) cap);
}

extern "C" cl_bool clo_sort_b200_debug(CloSort* sorter, CCLQueue* cq, cl_ulong out[18]) {
	if (!sorter || !cq || !out) return CL_FALSE;
	CloDeviceGuard g(cq->ctx->dev.ordinal);
	unsigned long long tmp[18];
	if (clo_radix_debug(sorter->rs, cq->stream, tmp) != 0) return CL_FALSE;
	for (int i = 0; i < 18; ++i) out[i] = tmp[i];
	return CL_TRUE;
}


extern "C" CCLEvent* clo_sort_pairs_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
		CCLBuffer* keys, CCLBuffer* payload, size_t numel, GError** err) {
	if (!sorter || (err && *err) || !cq_exec) return NULL;
	const size_t ks_bytes = clo_type_sizeof(sorter->elem_type);
	if (!keys || !payload || keys->size < numel * ks_bytes || payload->size < numel * sizeof(cl_uint)) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "sort pairs: buffers too small for %zu pairs", numel);
		return NULL;
	}
	if (ks_bytes != 4 && ks_bytes != 8) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "sort pairs: keys must be 4 or 8 bytes wide");
		return NULL;
	}
	if (sorter->ks.key_kind == CLO_KIND_FLOAT) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "sort pairs: integer keys only");
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_radix_onesweep_pairs");
	const char* msg = NULL;
	cudaError_t rc = clo_radix_sort(sorter->rs, clo_sm_count(cq_exec->ctx->dev.ordinal), ks_bytes, sorter->ks,
		(uint32_t) (8 * ks_bytes), keys->ptr, keys->ptr, (const uint32_t*) payload->ptr, (uint32_t*) payload->ptr,
		numel, cq_exec->stream, &msg);
	clo_queue_end(cq_exec, evt);
	if (rc != cudaSuccess && msg) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg); return NULL; }
	if (clo_cuda_failed(rc, err, "clo_radix_sort (pairs)")) return NULL;
	return evt;
}

extern "C" CCLEvent* clo_sort_partition_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
		CCLBuffer* keys_in, CCLBuffer* payload_in, CCLBuffer* keys_out, CCLBuffer* payload_out,
		size_t numel, cl_ulong gidx0, CCLBuffer* splitter_keys, CCLBuffer* splitter_idx, cl_uint nparts,
		CCLBuffer* counts_out, GError** err) {
	if (!sorter || (err && *err) || !cq_exec) return NULL;
	const size_t kb = clo_type_sizeof(sorter->elem_type);
	if (!keys_in || !keys_out || !counts_out || keys_in->size < numel * kb || keys_out->size < numel * kb ||
			counts_out->size < nparts * sizeof(cl_ulong) ||
			(nparts > 1 && (!splitter_keys || !splitter_idx || splitter_keys->size < (nparts - 1) * kb ||
				splitter_idx->size < (nparts - 1) * sizeof(cl_ulong))) ||
			((payload_in == NULL) != (payload_out == NULL))) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "partition: invalid buffers");
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_radix_partition");
	const char* msg = NULL;
	cudaError_t rc = clo_radix_partition(sorter->rs, kb, keys_in->ptr,
		payload_in ? (const uint32_t*) payload_in->ptr : NULL, keys_out->ptr,
		payload_out ? (uint32_t*) payload_out->ptr : NULL, numel, gidx0,
		splitter_keys ? splitter_keys->ptr : NULL, splitter_idx ? (const uint64_t*) splitter_idx->ptr : NULL,
		nparts, (uint64_t*) counts_out->ptr, cq_exec->stream, &msg);
	clo_queue_end(cq_exec, evt);
	if (rc != cudaSuccess && msg) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg); return NULL; }
	if (clo_cuda_failed(rc, err, "clo_radix_partition")) return NULL;
	return evt;
}

/* Sample sort, fused partition + exchange: stage 1 (bucket sizes) and stage 2 (scatter into
 * the receive buffers of the destination ranks).  dest_ptrs / payload_dest_ptrs are device
 * arrays of nparts raw device addresses (own buffer or peer memory imported with
 * clo_b200_ipc_import); first_slot[q] is the element index in destination q at which this
 * rank's bucket starts; *ok_flag == 0 makes the scatter a no-op (receive buffer too small). */
extern "C" CCLEvent* clo_sort_partition_count_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
		CCLBuffer* keys_in, size_t numel, cl_ulong gidx0, CCLBuffer* splitter_keys, CCLBuffer* splitter_idx,
		cl_uint nparts, CCLBuffer* counts_out, GError** err) {
	if (!sorter || (err && *err) || !cq_exec) return NULL;
	const size_t kb = clo_type_sizeof(sorter->elem_type);
	if (!keys_in || !counts_out || keys_in->size < numel * kb || counts_out->size < nparts * sizeof(cl_ulong) ||
			(nparts > 1 && (!splitter_keys || !splitter_idx || splitter_keys->size < (nparts - 1) * kb ||
				splitter_idx->size < (nparts - 1) * sizeof(cl_ulong)))) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "partition count: invalid buffers");
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_partition_count");
	const char* msg = NULL;
	cudaError_t rc = clo_radix_partition_count(sorter->rs, kb, keys_in->ptr, numel, gidx0,
		splitter_keys ? splitter_keys->ptr : NULL, splitter_idx ? (const uint64_t*) splitter_idx->ptr : NULL,
		nparts, (uint64_t*) counts_out->ptr, cq_exec->stream, &msg);
	clo_queue_end(cq_exec, evt);
	if (rc != cudaSuccess && msg) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg); return NULL; }
	if (clo_cuda_failed(rc, err, "clo_partition_count")) return NULL;
	return evt;
}

extern "C" CCLEvent* clo_sort_partition_scatter_with_device_data(CloSort* sorter, CCLQueue* cq_exec,
		CCLBuffer* keys_in, CCLBuffer* payload_in, size_t numel, cl_ulong gidx0, CCLBuffer* splitter_keys,
		CCLBuffer* splitter_idx, cl_uint nparts, CCLBuffer* first_slot, CCLBuffer* dest_ptrs,
		CCLBuffer* payload_dest_ptrs, CCLBuffer* ok_flag, GError** err) {
	if (!sorter || (err && *err) || !cq_exec) return NULL;
	const size_t kb = clo_type_sizeof(sorter->elem_type);
	if (!keys_in || !first_slot || !dest_ptrs || keys_in->size < numel * kb ||
			first_slot->size < nparts * sizeof(cl_ulong) || dest_ptrs->size < nparts * sizeof(void*) ||
			(payload_in && (!payload_dest_ptrs || payload_dest_ptrs->size < nparts * sizeof(void*))) ||
			(nparts > 1 && (!splitter_keys || !splitter_idx))) {
		g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "partition scatter: invalid buffers");
		return NULL;
	}
	CloDeviceGuard g(cq_exec->ctx->dev.ordinal);
	ccl_event* evt = clo_queue_begin(cq_exec, "clo_partition_scatter");
	const char* msg = NULL;
	cudaError_t rc = clo_radix_partition_scatter(sorter->rs, kb, keys_in->ptr,
		payload_in ? (const uint32_t*) payload_in->ptr : NULL, numel, gidx0,
		splitter_keys ? splitter_keys->ptr : NULL, splitter_idx ? (const uint64_t*) splitter_idx->ptr : NULL, nparts,
		(const uint64_t*) first_slot->ptr, (void* const*) dest_ptrs->ptr,
		payload_dest_ptrs ? (void* const*) payload_dest_ptrs->ptr : NULL, ok_flag ? (const int*) ok_flag->ptr : NULL,
		cq_exec->stream, &msg);
	clo_queue_end(cq_exec, evt);
	if (rc != cudaSuccess && msg) { g_set_error(err, CLO_ERROR, CLO_ERROR_ARGS, "%s", msg); return NULL; }
	if (clo_cuda_failed(rc, err, "clo_partition_scatter")) return NULL;
	return evt;
}
