#!/usr/bin/env python3
"""tests/golden/make_golden.py -- generate the golden fixtures in this directory by
running the REFERENCE'S OWN kernels on the CPU (oracle/_ref/libclo_ref.so, built by
oracle/build_ref.py from /root/reference).  Only runs where /root/reference is
mounted; the .npz files it writes are committed and are what the tests read.

Each fixture holds the inputs and the outputs of the reference kernels, driven by
the reference's enqueue sequences, for one family:
  ref_rng.npz   seeds after clo_rng_init (3 hashes), bulk output + final states
  ref_scan.npz  exclusive scans (6 type pairs, several n / lws)
  ref_sort.npz  sbitonic, gselect, satradix (radix 2/16/256) results
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

NP = {"uint": np.uint32, "int": np.int32, "ulong": np.uint64, "long": np.int64, "uchar": np.uint8,
      "ushort": np.uint16, "float": np.float32}
SEED_SIZE = {"lcg": 8, "xorshift64": 8, "xorshift128": 16, "mwc64x": 8, "parkmiller": 4, "tauslcg": 16}


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def main():
    if not os.path.exists(build_ref.OUT_LIB):
        build_ref.main()
    L = ctypes.CDLL(build_ref.OUT_LIB)
    rs = np.random.default_rng(20261018)

    # ------------------------------------------------------------------ rng
    out = {}
    G, runs, main_seed = 96, 6, 1234
    for rng in build_ref.RNGS:
        for h in build_ref.HASHES:
            st = np.zeros(G * SEED_SIZE[rng], dtype=np.uint8)
            f = getattr(L, "ref_rng_init_%s_%s" % (rng, h))
            f.argtypes = [ctypes.c_ulong, ctypes.c_void_p, ctypes.c_size_t]
            f(main_seed, ptr(st), G)
            out["%s/%s/seeds" % (rng, h)] = st.copy()
            bench = getattr(L, "ref_rng_bench_%s" % rng)
            bench.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                              ctypes.c_uint, ctypes.c_uint]
            res = np.zeros((runs, G), dtype=np.uint32)
            bench(ptr(st), ptr(res), G, runs, 32, 0)
            out["%s/%s/out32" % (rng, h)] = res.copy()
            out["%s/%s/states_after" % (rng, h)] = st.copy()
            res8 = np.zeros((3, G), dtype=np.uint32)
            bench(ptr(st), ptr(res8), G, 3, 8, 0)
            out["%s/%s/out8_cont" % (rng, h)] = res8.copy()
            resm = np.zeros((3, G), dtype=np.uint32)
            bench(ptr(st), ptr(resm), G, 3, 0, 1000)
            out["%s/%s/outmax1000_cont" % (rng, h)] = resm.copy()
    out["meta"] = np.array([G, runs, main_seed], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ref_rng.npz"), **out)

    # ----------------------------------------------------------------- scan
    out = {}
    for name, elem, sm in build_ref.SCAN_VARIANTS:
        f = getattr(L, "ref_scan_%s" % name)
        f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
        for n, lws in ((64, 8), (1024, 8), (1024, 32), (4096, 16), (8192, 64)):
            if elem == "float":
                a = rs.random(n).astype(np.float32)
            elif n == 1024 and lws == 32:
                info = np.iinfo(NP[elem])
                a = rs.integers(info.min, info.max, size=n, dtype=NP[elem], endpoint=True)  # wrap-around
            else:
                a = rs.integers(0, 128, size=n).astype(NP[elem])
            o = np.zeros(n, dtype=NP[sm])
            assert f(ptr(a), ptr(o), n, lws) == 0
            out["%s/%d/%d/in" % (name, n, lws)] = a
            out["%s/%d/%d/out" % (name, n, lws)] = o
    np.savez_compressed(os.path.join(HERE, "ref_scan.npz"), **out)

    # ----------------------------------------------------------------- sort
    out = {}
    for name, elem, key, compare, get_key in build_ref.SORT_VARIANTS:
        dt = NP[elem]

        def mk(n):
            if elem == "float":
                return ((rs.random(n) - 0.5) * 1000).astype(np.float32)
            info = np.iinfo(dt)
            a = rs.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
            if name == "ulong_keylo8":
                a = (a & ~np.uint64(0xFF)) | rs.integers(0, 5, size=n).astype(np.uint64)
            if name == "ulong_keyhi32":
                a = (a & np.uint64(0xFFFFFFFF)) | (rs.integers(0, 40, size=n).astype(np.uint64) << np.uint64(32))
            return a

        sb = getattr(L, "ref_sort_sbitonic_%s" % name)
        sb.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        for n in (16, 1024, 8192):
            a = mk(n)
            o = a.copy()
            assert sb(ptr(o), n) == 0
            out["sbitonic/%s/%d/in" % (name, n)] = a
            out["sbitonic/%s/%d/out" % (name, n)] = o
        gs = getattr(L, "ref_sort_gselect_%s" % name)
        gs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        a = mk(700)
        o = np.zeros_like(a)
        assert gs(ptr(a), ptr(o), a.size) == 0
        out["gselect/%s/700/in" % name] = a
        out["gselect/%s/700/out" % name] = o
        if (name, elem, key, compare, get_key) in build_ref.RADIX_VARIANTS:
            for nb in build_ref.RADIX_BITS:
                sr = getattr(L, "ref_sort_satradix%d_%s" % (nb, name))
                sr.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t]
                for n, lws in ((1024, 16), (2048, 256)):
                    if nb == 1 and n > 1024:
                        continue
                    a = mk(n)
                    o = a.copy()
                    assert sr(ptr(o), n, lws) == 0, (name, nb, n, lws)
                    out["satradix%d/%s/%d/%d/in" % (1 << nb, name, n, lws)] = a
                    out["satradix%d/%s/%d/%d/out" % (1 << nb, name, n, lws)] = o
    np.savez_compressed(os.path.join(HERE, "ref_sort.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
