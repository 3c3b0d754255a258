"""tools/sanitize_driver.py -- the small configurations compute-sanitizer is run on:
every kernel family once, through the C ABI with host data (no torch), results checked against the
oracle.  tools/sanitize.sh runs it under memcheck, racecheck and synccheck."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo  # noqa: E402
import oracle  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
ctx = clo.Context(0)
q = clo.Queue(ctx)
rng = np.random.default_rng(1)
if which in ("all", "sort"):
    a = rng.integers(0, 2**32, size=(1 << 16) + 123, dtype=np.uint64).astype(np.uint32)
    s = clo.CloSort("satradix", ctx, clo.UINT)
    assert np.array_equal(s.with_host_data(a, q), np.sort(a))           # histogram, bin scan, onesweep v6 + propagators
    s.destroy()
    a8 = rng.integers(0, 2**63, size=(1 << 15) + 7, dtype=np.uint64)
    s = clo.CloSort("satradix", ctx, clo.ULONG)
    assert np.array_equal(s.with_host_data(a8, q), np.sort(a8))
    s.destroy()
    b = rng.integers(0, 2**32, size=5000, dtype=np.uint64).astype(np.uint32)
    for alg in ("sbitonic", "gselect"):
        s = clo.CloSort(alg, ctx, clo.UINT)
        assert np.array_equal(s.with_host_data(b, q), np.sort(b))
        s.destroy()
    print("sort ok", flush=True)
if which in ("all", "scan"):
    for n in ((1 << 22), (1 << 22) + 77, 100003):                      # copy-engine kernel, cp.async ring, look-back
        c = rng.integers(0, 1000, size=n, dtype=np.uint32)
        sc = clo.CloScan("blelloch", ctx, clo.UINT, clo.UINT)
        want = (np.cumsum(c, dtype=np.uint64) - c).astype(np.uint32)
        assert np.array_equal(sc.with_host_data(c, q), want)
        sc.destroy()
    f = rng.random(1 << 22, dtype=np.float32)
    sc = clo.CloScan("blelloch", ctx, clo.FLOAT, clo.FLOAT)
    got = sc.with_host_data(f, q)
    ref = np.cumsum(f.astype(np.float64)) - f
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-3)
    sc.destroy()
    print("scan ok", flush=True)
if which in ("all", "rng"):
    for name in oracle.RNG_NAMES:
        r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, 4096, 7, "KNUTH(x)", q)
        got = r.generate_host(4, queue=q)
        want, _ = oracle.rng_generate(name, oracle.rng_seeds_dev_gid(name, 1, 7, 4096), 4096, 4)
        assert np.array_equal(got, want)
        r.destroy()
    print("rng ok", flush=True)
q.destroy()
ctx.destroy()
