"""CPU: the C-ABI library loads and exports every symbol the public headers declare
(include/*.h, include/cl_ops/*.h, include/compat/*.h).  No compute call is made."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cl_ops_b200", "libcl_ops.so")

DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?\b((?:clo|ccl|g)_\w+)\s*\(", re.M)
EXTERN_DATA = re.compile(r"^extern\s+const\s+[\w\s]+?\b(clo_\w+)\s*(?:\[\])?;", re.M)


def declared_symbols():
    funcs, data = set(), set()
    for path in glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)            # comments
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)             # preprocessor lines
        text = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)  # vtables
        for m in DECL.finditer(text):
            funcs.add(m.group(1))
        for m in EXTERN_DATA.finditer(text):
            data.add(m.group(1))
    return funcs, data


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.fail("libcl_ops.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    return ctypes.CDLL(LIB)


def test_headers_declare_the_reference_api():
    funcs, data = declared_symbols()
    for name in ("clo_sort_new", "clo_sort_destroy", "clo_sort_with_host_data", "clo_sort_with_device_data",
                 "clo_scan_new", "clo_scan_destroy", "clo_scan_with_host_data", "clo_scan_with_device_data",
                 "clo_rng_new", "clo_rng_destroy", "clo_rng_get_source", "clo_rng_get_device_seeds",
                 "clo_rng_get_size", "clo_type_by_name", "clo_type_sizeof", "clo_type_get_name", "clo_nlpo2",
                 "clo_error_quark", "clo_sort_get_num_kernels", "clo_scan_get_sum_size"):
        assert name in funcs, name
    for name in ("clo_sort_sbitonic_def", "clo_sort_abitonic_def", "clo_sort_gselect_def",
                 "clo_sort_satradix_def", "clo_scan_blelloch_def", "clo_rng_infos"):
        assert name in data, name
    assert len(funcs) > 90


def test_library_exports_every_declared_symbol(lib):
    funcs, data = declared_symbols()
    missing = [s for s in sorted(funcs | data) if not hasattr(lib, s)]
    assert not missing, "declared in include/ but not exported: %s" % missing


def test_host_only_helpers_match_reference_semantics(lib):
    """clo_common helpers are pure host code (clo_common.c:54-223): callable without a GPU."""
    lib.clo_type_get_name.restype = ctypes.c_char_p
    lib.clo_type_sizeof.restype = ctypes.c_size_t
    names = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong", "half", "float", "double"]
    sizes = [1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8]
    for i, (nm, sz) in enumerate(zip(names, sizes)):
        assert lib.clo_type_get_name(i) == nm.encode()
        assert lib.clo_type_sizeof(i) == sz
        assert lib.clo_type_by_name(nm.encode(), None) == i
    assert lib.clo_type_get_name(11) is None and lib.clo_type_sizeof(-1) == 0
    lib.clo_nlpo2.restype = ctypes.c_uint
    for x, want in ((0, 0), (1, 1), (2, 2), (3, 4), (5, 8), (1000, 1024), (1 << 20, 1 << 20), ((1 << 20) + 1, 1 << 21)):
        assert lib.clo_nlpo2(x) == want
    assert lib.clo_ones32(0xF0F0) == 8 and lib.clo_tzc(16) == 4 and lib.clo_tzc(256) == 8
    assert lib.clo_sum(0) == 0 and lib.clo_sum(10) == 55

    class GError(ctypes.Structure):
        _fields_ = [("domain", ctypes.c_uint32), ("code", ctypes.c_int), ("message", ctypes.c_char_p)]
    err = ctypes.POINTER(GError)()
    assert lib.clo_type_by_name(b"quaternion", ctypes.byref(err)) == -1
    assert err and err.contents.code == 6 and b"quaternion" in err.contents.message   # CLO_ERROR_UNKNOWN_TYPE
    lib.clo_error_quark.restype = ctypes.c_uint32
    assert err.contents.domain == lib.clo_error_quark()
    lib.clo_b200_error_free(err)


def test_rng_info_table(lib):
    class Info(ctypes.Structure):
        _fields_ = [("name", ctypes.c_char_p), ("src", ctypes.c_char_p), ("seed_size", ctypes.c_size_t)]
    table = (Info * 7).in_dll(lib, "clo_rng_infos")
    got = [(t.name.decode(), t.seed_size) for t in table[:6]]
    assert got == [("lcg", 8), ("xorshift64", 8), ("xorshift128", 16), ("mwc64x", 8), ("parkmiller", 4), ("tauslcg", 16)]
    assert table[6].name is None
    assert all(b"clo_rng_next" in t.src for t in table[:6])


def test_product_has_no_cpu_fallback():
    """the package must not import, link or call the oracle"""
    for path in glob.glob(os.path.join(ROOT, "cl_ops_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
            text = open(path, errors="ignore").read()
            assert "import oracle" not in text and "liboracle" not in text and "clo_oracle" not in text, path


@pytest.mark.parametrize("header", ["clo_sort_abitonic.h", "clo_sort_sbitonic.h", "clo_sort_gselect.h",
                                    "clo_sort_satradix.h", "clo_scan_blelloch.h"])
def test_per_algorithm_headers_compile_on_their_own(header, tmp_path):
    """cl_ops.h:29-52 lists one public header per algorithm; a client may include one directly."""
    import subprocess
    src = tmp_path / "client.c"
    macro = {"clo_sort_abitonic.h": "CLO_SORT_ABITONIC_NUM_KERNELS", "clo_sort_sbitonic.h": "CLO_SORT_SBITONIC_NUM_KERNELS",
             "clo_sort_gselect.h": "CLO_SORT_GSELECT_NUM_KERNELS", "clo_sort_satradix.h": "CLO_SORT_SATRADIX_NUM_KERNELS",
             "clo_scan_blelloch.h": "CLO_SCAN_BLELLOCH_NUM_KERNELS"}[header]
    src.write_text("#include <cl_ops/%s>\nint n_kernels(void) { return %s; }\n" % (header, macro))
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "include", "compat"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_dist_rng_partition_is_a_partition():
    """clo_dist_rng_partition (pure host arithmetic, no device needed): contiguous, disjoint,
    complete, sizes differ by at most one -- for every world size the library supports."""
    import ctypes
    import cl_ops_b200 as clo
    for total in (0, 1, 7, 1 << 22, (1 << 32) + 12345):
        for world in range(1, 17):
            spans = [clo.CloDist.rng_partition(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert spans[-1][0] + spans[-1][1] == total
            sizes = [c for _, c in spans]
            assert max(sizes) - min(sizes) <= 1
