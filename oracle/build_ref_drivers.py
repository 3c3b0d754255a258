"""oracle/build_ref_drivers.py -- compile the reference's OWN benchmark drivers, unchanged, from
where they lie under /root/reference (src/benchmarks/clo_sort_bench.c, clo_scan_bench.c,
clo_rng_bench.c, clo_bench.c, src/tests/test_rng.c; the error macros of
src/cl_ops/common/_g_err_macros.h) against this repository's
headers (include/, include/compat/{glib,cf4ocl2}.h) and link them with cl_ops_b200/libcl_ops.so.

Outputs go to oracle/_ref/ only (git-ignored binaries; no reference source is copied).  This is
the "existing src/benchmarks drivers relink unchanged" check of BASELINE.json's north_star:
the drivers verify their own results (sorted order, clo_sort_bench.c:216-226; scan == serial
host scan, clo_scan_bench.c:253-270), so running them on the GPU box is a parity test written
by the reference's author.  clo_rng_bench and test_rng compile an OpenCL C kernel string through
ccl_program_*: the library's NVRTC-backed program / kernel shim (csrc/jit.cu) builds it at run time.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "oracle", "_ref")


def main():
    bench = os.path.join(REF, "src", "benchmarks")
    if not os.path.isdir(bench):
        print("build_ref_drivers: %s not present, nothing built" % bench)
        return 0
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "compat"),
           "-I" + bench, "-I" + os.path.join(REF, "src", "cl_ops")]
    objs = {}
    for name in ("clo_bench", "clo_sort_bench", "clo_scan_bench"):
        o = os.path.join(OUT, name + ".o")
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-w"] + inc + ["-c", os.path.join(bench, name + ".c"), "-o", o])
        objs[name] = o
    # clo_rng_bench: its header is a CMake template holding the bench kernel as a C string
    # (clo_rng_bench.in.h: CLO_RNG_BENCHMARK_SRC "@RNG_BENCHMARK_SRC@"); do CMake's configure step
    # here, into the build directory, from the .cl where it lies
    gen = os.path.join(OUT, "gen")
    os.makedirs(gen, exist_ok=True)
    cl = open(os.path.join(bench, "clo_rng_bench.cl")).read()
    esc = cl.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n\" \\\n\"")
    tmpl = open(os.path.join(bench, "clo_rng_bench.in.h")).read()
    with open(os.path.join(gen, "clo_rng_bench.h"), "w") as f:
        f.write(tmpl.replace("@RNG_BENCHMARK_SRC@", esc))
    o = os.path.join(OUT, "clo_rng_bench.o")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-w", "-I" + gen] + inc + ["-c", os.path.join(bench, "clo_rng_bench.c"), "-o", o])
    objs["clo_rng_bench"] = o
    # the reference's only unit test
    o = os.path.join(OUT, "test_rng.o")
    subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-w"] + inc + ["-c", os.path.join(REF, "src", "tests", "test_rng.c"), "-o", o])
    subprocess.check_call(["gcc", "-o", os.path.join(OUT, "test_rng"), o, "-L" + os.path.join(ROOT, "cl_ops_b200"), "-lcl_ops", "-lm",
                           "-Wl,-rpath,$ORIGIN/../../cl_ops_b200"])
    os.remove(o)
    for exe in ("clo_sort_bench", "clo_scan_bench", "clo_rng_bench"):
        subprocess.check_call(["gcc", "-o", os.path.join(OUT, exe), objs[exe], objs["clo_bench"],
                               "-L" + os.path.join(ROOT, "cl_ops_b200"), "-lcl_ops", "-lm",
                               "-Wl,-rpath,$ORIGIN/../../cl_ops_b200"])
    for o in objs.values():
        os.remove(o)
    print("build_ref_drivers: wrote %s/{clo_sort_bench,clo_scan_bench,clo_rng_bench,test_rng}" % OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
