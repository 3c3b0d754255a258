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


def test_reference_test_rng_unchanged():
    """src/tests/test_rng.c, the reference's only unit test: for every generator and every seeding
    mode it concatenates clo_rng_get_source() with an OpenCL C kernel, builds it (here: NVRTC),
    runs it and asserts no error and no leaked wrapper."""
    out = _run("test_rng")
    for path in ("/rng/seed-dev-gid", "/rng/seed-host-mt", "/rng/seed-ext-dev", "/rng/seed-ext-host"):
        assert path + ": OK" in out


@pytest.mark.parametrize("rng", ["lcg", "xorshift64", "xorshift128", "mwc64x", "parkmiller", "tauslcg"])
@pytest.mark.parametrize("hash_", ["KNUTH(x)", "XS1(x)"])
def test_reference_rng_bench_unchanged_stream_is_bit_exact(rng, hash_):
    """src/benchmarks/clo_rng_bench.c with --output stdout-uint: the numbers its own kernel
    (clo_rng_bench.cl, built at run time on top of clo_rng_get_source()) prints are the oracle's
    stream, run after run."""
    import numpy as np
    import oracle
    G, runs, seed = 2048, 5, 77
    out = _run("clo_rng_bench", "-r", rng, "-o", "stdout-uint", "-g", str(G), "-n", str(runs), "-s", str(seed),
               "--gid-hash", hash_, "-b", "32")
    got = np.array([int(t) for t in out.split()], dtype=np.uint64).astype(np.uint32)
    seeds = oracle.rng_seeds_dev_gid(rng, oracle.HASH_IDS[hash_], seed, G)
    want, _ = oracle.rng_generate(rng, seeds, G, runs)
    assert got.size == G * runs
    assert np.array_equal(got, want.reshape(-1))


def test_reference_rng_bench_unchanged_maxint_and_bits():
    import numpy as np
    import oracle
    G, runs = 1024, 3
    out = _run("clo_rng_bench", "-r", "mwc64x", "-o", "stdout-uint", "-g", str(G), "-n", str(runs), "-s", "5", "--gid-hash", "KNUTH(x)", "-m", "1000")
    got = np.array([int(t) for t in out.split()], dtype=np.uint32)
    seeds = oracle.rng_seeds_dev_gid("mwc64x", 1, 5, G)
    want, _ = oracle.rng_generate("mwc64x", seeds, G, runs, maxint=1000)
    assert np.array_equal(got, want.reshape(-1))
    out = _run("clo_rng_bench", "-r", "xorshift64", "-o", "stdout-uint", "-g", str(G), "-n", str(runs), "-s", "5", "--gid-hash", "KNUTH(x)", "-b", "8")
    got = np.array([int(t) for t in out.split()], dtype=np.uint32)
    seeds = oracle.rng_seeds_dev_gid("xorshift64", 1, 5, G)
    want, _ = oracle.rng_generate("xorshift64", seeds, G, runs, bits=8)
    assert np.array_equal(got, want.reshape(-1))


def _run_in(cwd, exe, *args):
    path = os.path.join(REF, exe)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time)" % exe)
    r = subprocess.run([path] + list(args), capture_output=True, timeout=600, cwd=cwd)
    assert r.returncode == 0, r.stdout[-1500:].decode(errors="replace") + r.stderr[-1500:].decode(errors="replace")
    return r.stdout


def test_reference_rng_bench_output_formats(tmp_path):
    """The wire / on-disk formats either side of the RNG path (clo_rng_bench.c:204-274, 314-323;
    read back by scripts/clo_rng_plot.py:33-57 and by dieharder): the unchanged driver writes
    file-tsv (one run per line, tab separated, np.loadtxt-able), file-dh (dieharder header
    "type: d / count / numbit", one number per line) and stdout-bin (raw little-endian uint32),
    and all three carry the oracle's stream."""
    import numpy as np
    import oracle
    rng, G, runs, seed, hash_ = "xorshift128", 512, 4, 9, "KNUTH(x)"
    seeds = oracle.rng_seeds_dev_gid(rng, oracle.HASH_IDS[hash_], seed, G)
    want, _ = oracle.rng_generate(rng, seeds, G, runs)
    want = np.asarray(want).reshape(runs, G)
    common = ["-r", rng, "-g", str(G), "-n", str(runs), "-s", str(seed), "--gid-hash", hash_, "-b", "32"]
    # file-tsv: out_<rng>_gid_<hash>.tsv in the working directory
    _run_in(str(tmp_path), "clo_rng_bench", "-o", "file-tsv", *common)
    tsv = tmp_path / ("out_%s_gid_%s.tsv" % (rng, hash_))
    assert tsv.exists()
    img = np.loadtxt(str(tsv), dtype=np.uint32)                  # what clo_rng_plot.py does
    assert img.shape == (runs, G) and np.array_equal(img, want)
    assert tsv.read_text().splitlines()[0].count("\t") == G       # "%u\t" per value, "\n" per run
    # file-dh: dieharder ASCII input
    _run_in(str(tmp_path), "clo_rng_bench", "-o", "file-dh", *common)
    dh = (tmp_path / ("out_%s_gid_%s.dh.txt" % (rng, hash_))).read_text().splitlines()
    assert dh[0] == "type: d" and dh[1] == "count: %d" % (G * runs) and dh[2] == "numbit: 32"
    assert np.array_equal(np.array([int(x) for x in dh[3:]], dtype=np.uint64).astype(np.uint32), want.reshape(-1))
    # stdout-bin: raw words
    raw = _run_in(str(tmp_path), "clo_rng_bench", "-o", "stdout-bin", *common)
    assert len(raw) == 4 * G * runs
    assert np.array_equal(np.frombuffer(raw, dtype="<u4"), want.reshape(-1))
