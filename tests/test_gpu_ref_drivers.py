"""The reference's OWN benchmark drivers (src/benchmarks/clo_sort_bench.c, clo_scan_bench.c),
compiled unchanged by oracle/build_ref_drivers.py and linked against libcl_ops.so, run on the
GPU: they check their results themselves (sorted order; scan == serial host scan)."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(exe, *args):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time)" % exe)
    r = subprocess.run([path] + list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r.stdout


@pytest.mark.parametrize("alg,typ,maxpo2", [("satradix", "uint", 22), ("satradix", "ulong", 20), ("satradix", "uchar", 18),
                                            ("sbitonic", "uint", 20), ("sbitonic", "float", 18), ("sbitonic", "long", 16),
                                            ("abitonic", "int", 18), ("gselect", "uint", 12)])
def test_reference_sort_bench_unchanged(alg, typ, maxpo2):
    out = _run("clo_sort_bench", "-a", alg, "-t", typ, "-n", str(maxpo2), "-r", "2", "-s", "3")
    lines = [l for l in out.splitlines() if "Mkeys/s" in l]
    assert len(lines) == maxpo2 - 3                      # one line per size 2^4 .. 2^maxpo2
    assert "did not work" not in out


@pytest.mark.parametrize("typ,sumtyp", [("uint", "ulong"), ("uint", "uint"), ("uchar", "uint"), ("ushort", "ulong")])
def test_reference_scan_bench_unchanged(typ, sumtyp):
    out = _run("clo_scan_bench", "-t", typ, "-y", sumtyp, "-i", "1000", "-n", "11", "-r", "2")
    lines = [l for l in out.splitlines() if "MValues/s" in l]
    assert len(lines) > 0
    for l in lines:
        assert "did not work" not in l and "Unverified" not in l
