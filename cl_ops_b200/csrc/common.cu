/*
 * common.cu -- CloType table, bit helpers, error quark.
 * Follows /root/reference/src/cl_ops/common/clo_common.c:54-223 (behaviour),
 * written against the compat GLib subset.
 */
#include "clo_internal.h"

#include <cstring>

namespace {
struct TypeInfo { const char* name; size_t size; };
/* index == CloType value (clo_common.in.h:108-120) */
const TypeInfo kTypes[] = {
	{"char", 1}, {"uchar", 1}, {"short", 2}, {"ushort", 2}, {"int", 4}, {"uint", 4},
	{"long", 8}, {"ulong", 8}, {"half", 2}, {"float", 4}, {"double", 8},
};
const int kNumTypes = sizeof(kTypes) / sizeof(kTypes[0]);
inline bool valid(int t) { return t >= 0 && t < kNumTypes; }
}

extern "C" const char* clo_type_get_name(CloType type) {
	return valid((int) type) ? kTypes[type].name : NULL;
}

extern "C" size_t clo_type_sizeof(CloType type) {
	return valid((int) type) ? kTypes[type].size : 0;
}

extern "C" CloType clo_type_by_name(const char* name, GError** err) {
	if (name)
		for (int i = 0; i < kNumTypes; ++i)
			if (strcmp(name, kTypes[i].name) == 0) return (CloType) i;
	g_set_error(err, CLO_ERROR, CLO_ERROR_UNKNOWN_TYPE, "Unknown type '%s'", name ? name : "(null)");
	return (CloType) -1;
}

extern "C" unsigned int clo_nlpo2(unsigned int x) {
	if ((x & (x - 1)) == 0) return x;
	x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
	return x + 1;
}

extern "C" unsigned int clo_ones32(unsigned int x) {
	return (unsigned int) __builtin_popcount(x);
}

extern "C" unsigned int clo_tzc(int x) {
	return clo_ones32((unsigned int) ((x & -x) - 1));
}

extern "C" unsigned int clo_sum(unsigned int x) {
	/* 0 + 1 + ... + x, modulo 2^32 like the recursive original */
	return (unsigned int) (((unsigned long long) x * ((unsigned long long) x + 1)) / 2);
}

extern "C" void clo_print_to_null(const gchar* string) { (void) string; }

extern "C" GQuark clo_error_quark(void) {
	return g_quark_from_static_string("clo-error-quark");
}

extern "C" const char* clo_b200_version(void) { return "cl_ops-b200 0.1.0 (sm_100a)"; }

extern "C" cl_ulong clo_b200_launch_count(void) {
	return clo_launches.load(std::memory_order_relaxed);
}
