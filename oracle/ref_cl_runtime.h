/*
 * oracle/ref_cl_runtime.h -- just enough of the OpenCL C execution model to run
 * the reference's own kernels (/root/reference/src/cl_ops/ **.cl) on the CPU:
 * address-space qualifiers, vector types, work-item functions, barrier().
 *
 * TEST INFRASTRUCTURE ONLY.  Used by oracle/build_ref.py, which pipes the
 * reference's .cl text (never copied into this repository) through g++ with this
 * header prepended.  A work-group is run as one OS thread per work-item so that
 * barrier() has its real semantics; kernels without barriers run serially.
 */
#ifndef CLO_REF_CL_RUNTIME_H
#define CLO_REF_CL_RUNTIME_H

#include <climits>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <barrier>
#include <functional>
#include <thread>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long ulong;   /* LP64: 64 bits, like OpenCL C's ulong */

#define __kernel
#define __global
#define __local            /* pointer parameters; in-kernel __local arrays are rewritten to `static` */
#define __constant const
#define CLK_LOCAL_MEM_FENCE 1
#define CLK_GLOBAL_MEM_FENCE 2

struct uint2 {
	union { struct { uint x, y; }; struct { uint s0, s1; }; };
	uint2() : x(0), y(0) {}
	uint2(uint a, uint b) : x(a), y(b) {}
};
struct uint4 {
	union { struct { uint x, y, z, w; }; struct { uint s0, s1, s2, s3; }; };
	uint4() : x(0), y(0), z(0), w(0) {}
	uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
};
struct uint8 {
	uint s0, s1, s2, s3, s4, s5, s6, s7;
	uint8() : s0(0), s1(0), s2(0), s3(0), s4(0), s5(0), s6(0), s7(0) {}
	uint8(uint a, uint b, uint c, uint d, uint e, uint f, uint g, uint h)
		: s0(a), s1(b), s2(c), s3(d), s4(e), s5(f), s6(g), s7(h) {}
};
struct ulong2 {
	ulong x, y;
	ulong2() : x(0), y(0) {}
	ulong2(ulong a, ulong b) : x(a), y(b) {}
};

static inline uint2 as_uint2(ulong v) { uint2 r; std::memcpy((void*) &r, &v, 8); return r; }
static inline uint4 as_uint4(ulong2 v) { uint4 r; std::memcpy((void*) &r, &v, 16); return r; }
static inline int as_int(uint v) { int r; std::memcpy(&r, &v, 4); return r; }
static inline uint as_uint(int v) { uint r; std::memcpy(&r, &v, 4); return r; }
static inline uint convert_uint(ulong v) { return (uint) v; }
static inline uint mul_hi(uint a, uint b) { return (uint) (((ulong) a * (ulong) b) >> 32); }

struct clo_ref_workitem {
	size_t gid, lid, lsz, grp, ngrp, gsz;
	std::barrier<>* bar;
};
static thread_local clo_ref_workitem clo_ref_wi;

static inline size_t get_global_id(uint) { return clo_ref_wi.gid; }
static inline size_t get_local_id(uint) { return clo_ref_wi.lid; }
static inline size_t get_local_size(uint) { return clo_ref_wi.lsz; }
static inline size_t get_group_id(uint) { return clo_ref_wi.grp; }
static inline size_t get_num_groups(uint) { return clo_ref_wi.ngrp; }
static inline size_t get_global_size(uint) { return clo_ref_wi.gsz; }
static inline void barrier(int) { if (clo_ref_wi.bar) clo_ref_wi.bar->arrive_and_wait(); }

/* NDRange without barriers: one thread walks every work-item. */
static inline void clo_ref_run_serial(size_t gws, size_t lws, const std::function<void()>& body) {
	if (lws == 0) lws = 1;
	for (size_t g = 0; g < gws; ++g) {
		clo_ref_wi = clo_ref_workitem{ g, g % lws, lws, g / lws, (gws + lws - 1) / lws, gws, nullptr };
		body();
	}
}

/* NDRange with barriers: work-groups one after the other, one thread per
 * work-item.  `per_group` is called before each group (to reset __local
 * memory the host passed as kernel arguments, if it wants to). */
static inline void clo_ref_run_groups(size_t gws, size_t lws, const std::function<void()>& body) {
	const size_t ngrp = gws / lws;
	for (size_t grp = 0; grp < ngrp; ++grp) {
		std::barrier<> bar((std::ptrdiff_t) lws);
		std::vector<std::thread> th;
		th.reserve(lws);
		for (size_t lid = 0; lid < lws; ++lid) {
			th.emplace_back([&, lid]() {
				clo_ref_wi = clo_ref_workitem{ grp * lws + lid, lid, lws, grp, ngrp, gws, &bar };
				body();
			});
		}
		for (auto& t : th) t.join();
	}
}

#endif
