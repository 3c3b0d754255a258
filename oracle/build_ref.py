#!/usr/bin/env python3
"""oracle/build_ref.py -- compile the REFERENCE'S OWN OpenCL C kernels for the CPU.

TEST INFRASTRUCTURE ONLY.  The .cl files are read where they lie under
/root/reference/src (they are never copied into this repository), passed
through three mechanical rewrites that turn OpenCL C into C++ --

  (T)(a, b, ...) vector literals      ->  T(a, b, ...)         (constructor call)
  `__local T name[N];` inside a kernel ->  `static T name[N];` (one group at a time)
  nothing else

-- wrapped in one namespace per macro configuration (the reference injects
element/key/compare/hash macros at JIT time: clo_sort_abstract.c:144-168,
clo_scan_abstract.c:122-125, clo_rng.c:101-109), and piped to g++ together with
oracle/ref_cl_runtime.h (work-items, barrier) and host loops that restate the
reference's enqueue sequences (cited below).  Output: oracle/_ref/libclo_ref.so
only.  The reference's host code itself needs cf4ocl2 + GLib + an OpenCL runtime,
none of which exist in this image, so it is "unbuildable" as shipped.

Usage: python oracle/build_ref.py [--reference /root/reference] [--emit]
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT_LIB = os.path.join(OUT_DIR, "libclo_ref.so")

VEC_LIT = re.compile(r"\((uint2|uint4|uint8|ulong2|clo_statetype)\)\s*\(")
LOCAL_DECL = re.compile(r"^(\s*)__local\s+([^;(),]+;)\s*$")

RNGS = ["lcg", "xorshift64", "xorshift128", "mwc64x", "parkmiller", "tauslcg"]
HASHES = {"none": None, "knuth": "KNUTH(x)", "xs1": "XS1(x)"}

# (name, elem, key, compare, get_key)   -- clo_sort_abstract.c:144-168
SORT_VARIANTS = [
    ("uint", "uint", "uint", "((a) > (b))", "(x)"),
    ("uint_desc", "uint", "uint", "((a) < (b))", "(x)"),
    ("int", "int", "int", "((a) > (b))", "(x)"),
    ("ulong", "ulong", "ulong", "((a) > (b))", "(x)"),
    ("uchar", "uchar", "uchar", "((a) > (b))", "(x)"),
    ("ushort", "ushort", "ushort", "((a) > (b))", "(x)"),
    ("float", "float", "float", "((a) > (b))", "(x)"),
    ("ulong_keylo8", "ulong", "uchar", "((a) > (b))", "((x) & 0xFF)"),
    ("ulong_keyhi32", "ulong", "uint", "((a) > (b))", "((x) >> 32)"),
]
# satradix digit widths compiled per variant (CLO_SORT_NUM_BITS, clo_sort_satradix.c:430-431)
RADIX_BITS = [4, 8, 1]
# integer-keyed variants only: `key >> b` does not compile for float (clo_sort_satradix.cl:61)
RADIX_VARIANTS = [v for v in SORT_VARIANTS if v[2] != "float" and "desc" not in v[0]]

# (name, elem, sum)   -- clo_scan_abstract.c:122-125
SCAN_VARIANTS = [
    ("uint_uint", "uint", "uint"), ("uint_ulong", "uint", "ulong"), ("int_long", "int", "long"),
    ("uchar_uint", "uchar", "uint"), ("float_float", "float", "float"), ("ulong_ulong", "ulong", "ulong"),
]


def cl(ref, rel):
    """One reference .cl file, rewritten line by line (see module docstring)."""
    out = []
    with open(os.path.join(ref, "src", rel)) as f:
        for line in f:
            line = VEC_LIT.sub(lambda m: m.group(1) + "(", line)
            m = LOCAL_DECL.match(line)
            if m:
                line = "%sstatic %s\n" % (m.group(1), m.group(2))
            out.append(line)
    return "".join(out)


def generate(ref):
    g = []
    w = g.append
    w('#include "ref_cl_runtime.h"\n#include <cstdlib>\n')
    w("static size_t clo_ref_log2(size_t x) { size_t l = 0; while (((size_t) 1 << l) < x) ++l; return l; }\n")

    # ---------------- RNG: generators, init kernel per hash, bench kernel ----------------
    for rng in RNGS:
        w("namespace ns_rng_%s {\n" % rng)
        w(cl(ref, "cl_ops/rng/clo_rng_workitem.cl"))
        w(cl(ref, "cl_ops/rng/clo_rng_%s.cl" % rng))
        w(cl(ref, "cl_ops/rng/clo_rng_api.cl"))
        for hname, hmacro in HASHES.items():
            w("namespace init_%s {\n" % hname)
            # clo_rng.c:101-109: "#define CLO_RNG_HASH(x) <hash>" prepended to the init source
            w("#define CLO_RNG_HASH(x) %s\n" % (hmacro if hmacro else "x"))
            w(cl(ref, "cl_ops/rng/clo_rng_init.cl"))
            w("#undef CLO_RNG_HASH\n#undef KNUTH\n#undef XS1\n}\n")
        w("namespace bench_bits {\n" + cl(ref, "benchmarks/clo_rng_bench.cl") + "}\n")
        w("namespace bench_maxint {\n#define CLO_RNG_BENCHMARK_MAXINT 1\n" +
          cl(ref, "benchmarks/clo_rng_bench.cl") + "#undef CLO_RNG_BENCHMARK_MAXINT\n}\n")
        w("#undef clo_ulong2statetype\n#undef GLOBAL_SIZE\n#undef GID1\n#undef GID2\n#undef GID4\n#undef GID8\n")
        w("}\n")
        for hname in HASHES:
            # clo_rng.c:125-127: clo_rng_init over seeds_count work-items
            w('extern "C" void ref_rng_init_%s_%s(unsigned long main_seed, void* seeds, size_t count) {\n'
              "\tclo_ref_run_serial(count, 1, [&]() { ns_rng_%s::init_%s::clo_rng_init(main_seed, "
              "(ns_rng_%s::clo_statetype*) seeds); });\n}\n" % (rng, hname, rng, hname, rng))
        # clo_rng_bench.c:302-324: `runs` launches of clo_rng_bench, row r of the output each
        w('extern "C" void ref_rng_bench_%s(void* seeds, unsigned* result, size_t G, size_t runs, '
          "unsigned bits, unsigned maxint) {\n"
          "\tfor (size_t r = 0; r < runs; ++r) {\n"
          "\t\tif (maxint) clo_ref_run_serial(G, 1, [&]() { ns_rng_%s::bench_maxint::clo_rng_bench("
          "(ns_rng_%s::clo_statetype*) seeds, result + r * G, maxint); });\n"
          "\t\telse clo_ref_run_serial(G, 1, [&]() { ns_rng_%s::bench_bits::clo_rng_bench("
          "(ns_rng_%s::clo_statetype*) seeds, result + r * G, bits); });\n"
          "\t}\n}\n" % (rng, rng, rng, rng, rng))

    # ---------------- scan ----------------
    for name, elem, sm in SCAN_VARIANTS:
        w("namespace ns_scan_%s {\n#define CLO_SCAN_ELEM_TYPE %s\n#define CLO_SCAN_SUM_TYPE %s\n" % (name, elem, sm))
        w("typedef %s selem_t; typedef %s sum_t;\n" % (elem, sm))
        w(cl(ref, "cl_ops/scan/clo_scan_blelloch.cl"))
        w("#undef CLO_SCAN_ELEM_TYPE\n#undef CLO_SCAN_SUM_TYPE\n}\n")
        # host sequence: clo_scan_blelloch.c:130-141 (sizes), :155-195 (three launches)
        w('extern "C" int ref_scan_%s(const void* in, void* out, size_t numel, size_t lws) {\n' % name)
        w("\tusing namespace ns_scan_%s;\n" % name)
        w("""	if (lws == 0 || numel < 2 * lws) return -1;
	size_t realws = numel / 2;
	size_t gws_wgscan = ((realws + lws - 1) / lws) * lws;
	if (gws_wgscan > lws * lws) gws_wgscan = lws * lws;
	size_t ws_wgsumsscan = (gws_wgscan / lws) / 2;
	size_t gws_addwgsums = ((numel + lws - 1) / lws) * lws;
	uint blocks_per_wg = (uint) ((numel / 2 + gws_wgscan - 1) / gws_wgscan);
	std::vector<sum_t> wgsums(gws_wgscan / lws + 2 * lws);
	std::vector<sum_t> aux(2 * lws);
	std::vector<sum_t> out_pad(gws_addwgsums);
	clo_ref_run_groups(gws_wgscan, lws, [&]() {
		workgroupScan((selem_t*) in, out_pad.data(), wgsums.data(), aux.data(), (uint) numel, blocks_per_wg); });
	if (gws_wgscan > lws) {
		clo_ref_run_groups(ws_wgsumsscan, ws_wgsumsscan, [&]() { workgroupSumsScan(wgsums.data(), aux.data()); });
		clo_ref_run_groups(gws_addwgsums, lws, [&]() { addWorkgroupSums(wgsums.data(), out_pad.data(), blocks_per_wg); });
	}
	std::memcpy(out, out_pad.data(), numel * sizeof(sum_t));
	return 0;
}
""")

    # ---------------- sort ----------------
    for name, elem, key, compare, get_key in SORT_VARIANTS:
        macros = ("#define CLO_SORT_ELEM_TYPE %s\n#define CLO_SORT_KEY_TYPE %s\n"
                  "#define CLO_SORT_COMPARE(a, b) %s\n#define CLO_SORT_KEY_GET(x) %s\n" % (elem, key, compare, get_key))
        unmac = "#undef CLO_SORT_ELEM_TYPE\n#undef CLO_SORT_KEY_TYPE\n#undef CLO_SORT_COMPARE\n#undef CLO_SORT_KEY_GET\n"
        w("namespace ns_sort_%s {\n%stypedef %s selem_t; typedef %s skey_t;\n" % (name, macros, elem, key))
        w("namespace sb {\n" + cl(ref, "cl_ops/sort/clo_sort_sbitonic.cl") + "}\n")
        w("namespace gs {\n" + cl(ref, "cl_ops/sort/clo_sort_gselect.cl") + "}\n")
        if (name, elem, key, compare, get_key) in RADIX_VARIANTS:
            for nb in RADIX_BITS:
                w("namespace sr%d {\n#define CLO_SORT_NUM_BITS %d\n" % (nb, nb))
                w(cl(ref, "cl_ops/sort/clo_sort_satradix.cl"))
                w("#undef CLO_SORT_NUM_BITS\n#undef CLO_SORT_RADIX\n#undef CLO_SORT_RADIX1\n}\n")
        w(unmac + "}\n")
        # sbitonic host loop: clo_sort_sbitonic.c:73-118 (gws = nlpo2(n)/2; stages; steps)
        w('extern "C" int ref_sort_sbitonic_%s(void* data, size_t numel) {\n' % name)
        w("""	if (numel < 2 || (numel & (numel - 1))) return -1;
	using namespace ns_sort_%s;
	size_t gws = numel / 2;
	uint tot_stages = (uint) clo_ref_log2(gws * 2);
	for (uint stage = 1; stage <= tot_stages; ++stage)
		for (uint step = stage; step > 0; --step)
			clo_ref_run_serial(gws, 1, [&]() { sb::sbitonic((selem_t*) data, stage, step); });
	return 0;
}
""" % name)
        # gselect: clo_sort_gselect.c:75-78,110 (gws = numel, one launch)
        w('extern "C" int ref_sort_gselect_%s(const void* in, void* out, size_t numel) {\n' % name)
        w("\tusing namespace ns_sort_%s;\n"
          "\tclo_ref_run_serial(numel, 1, [&]() { gs::gselect((selem_t*) in, (selem_t*) out, (ulong) numel); });\n"
          "\treturn 0;\n}\n" % name)
        if (name, elem, key, compare, get_key) in RADIX_VARIANTS:
            for nb in RADIX_BITS:
                # satradix host loop: clo_sort_satradix.c:166-169,184-200,242-313; the global scan of the
                # counters (clo_sort_satradix.c:298-299) is run through the reference scan kernels (uint,uint)
                w('extern "C" int ref_sort_satradix%d_%s(void* data, size_t numel, size_t lws) {\n' % (nb, name))
                w("""	using namespace ns_sort_%(name)s;
	const uint radix = 1u << %(nb)d;
	if (numel < 2 || (numel & (numel - 1))) return -1;
	if (lws < radix) lws = radix;
	if (lws > numel || (lws & (lws - 1))) return -1;
	const size_t numel_eff = numel;
	const size_t num_wgs = numel_eff / lws + numel_eff %% lws;
	const uint total_digits = (uint) (sizeof(selem_t) * 8 / %(nb)d);
	const uint array_len = (uint) (numel_eff / num_wgs);
	std::vector<selem_t> data_aux(numel_eff);
	std::vector<uint> offsets(num_wgs * radix), counters(num_wgs * radix), counters_sum(num_wgs * radix);
	std::vector<selem_t> data_local(array_len);
	std::vector<uint> scan_local(array_len), offsets_local(radix), counters_local(radix);
	std::vector<skey_t> digits_local(array_len);
	selem_t* d = (selem_t*) data;
	for (uint i = 0; i < total_digits; ++i) {
		uint start_bit = i * %(nb)d;
		clo_ref_run_groups(numel_eff, lws, [&]() {
			sr%(nb)d::satradix_localsort(d, data_aux.data(), data_local.data(), scan_local.data(), start_bit); });
		clo_ref_run_groups(numel_eff, lws, [&]() {
			sr%(nb)d::satradix_histogram(data_aux.data(), offsets.data(), counters.data(), offsets_local.data(),
				counters_local.data(), digits_local.data(), start_bit, array_len); });
		size_t cn = num_wgs * radix;
		size_t scan_lws = lws;
		while (cn < 2 * scan_lws) scan_lws /= 2;
		if (scan_lws == 0 || ref_scan_uint_uint(counters.data(), counters_sum.data(), cn, scan_lws)) {
			/* degenerate sizes the reference scan cannot take: serial exclusive scan */
			uint acc = 0;
			for (size_t k = 0; k < cn; ++k) { counters_sum[k] = acc; acc += counters[k]; }
		}
		clo_ref_run_groups(numel_eff, lws, [&]() {
			sr%(nb)d::satradix_scatter(d, data_aux.data(), offsets.data(), counters_sum.data(), data_local.data(),
				offsets_local.data(), counters_local.data(), start_bit); });
	}
	return 0;
}
""" % dict(name=name, nb=nb))
    return "".join(g)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--emit", action="store_true", help="print the generated C++ instead of compiling")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(args.reference, "src", "cl_ops")):
        print("build_ref: %s not present -- nothing to do (the prebuilt _ref/ travels with the repo)" % args.reference)
        return 0
    src = generate(args.reference)
    if args.emit:
        sys.stdout.write(src)
        return 0
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["/usr/bin/g++", "-x", "c++", "-std=c++20", "-O1", "-w", "-fPIC", "-shared", "-pthread",
           "-I", HERE, "-o", OUT_LIB, "-"]
    subprocess.run(cmd, input=src.encode(), check=True)
    print("build_ref: wrote", OUT_LIB)
    return 0


if __name__ == "__main__":
    sys.exit(main())
