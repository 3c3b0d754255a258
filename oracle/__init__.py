"""oracle -- ctypes loader for the CPU oracle (oracle/clo_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from cl_ops_b200 (the product).
See clo_oracle.h for the parity-pinning statement.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

# CloType ids (src/cl_ops/common/clo_common.in.h:108-120)
CHAR, UCHAR, SHORT, USHORT, INT, UINT, LONG, ULONG, HALF, FLOAT, DOUBLE = range(11)
TYPE_NAMES = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong",
              "half", "float", "double"]
NP_TYPES = {CHAR: np.int8, UCHAR: np.uint8, SHORT: np.int16, USHORT: np.uint16,
            INT: np.int32, UINT: np.uint32, LONG: np.int64, ULONG: np.uint64,
            FLOAT: np.float32, DOUBLE: np.float64}
# clo_rng_infos order (src/cl_ops/rng/clo_rng.c:60-68)
RNG_NAMES = ["lcg", "xorshift64", "xorshift128", "mwc64x", "parkmiller", "tauslcg"]
RNG_SEED_SIZE = {"lcg": 8, "xorshift64": 8, "xorshift128": 16, "mwc64x": 8,
                 "parkmiller": 4, "tauslcg": 16}
HASH_NONE, HASH_KNUTH, HASH_XS1 = 0, 1, 2
HASH_IDS = {None: 0, "": 0, "x": 0, "KNUTH(x)": 1, "XS1(x)": 2}


def build(force=False):
    """Compile liboracle.so (and nothing else) with the committed Makefile."""
    src = os.path.join(_HERE, "clo_oracle.c")
    if force or not os.path.exists(_LIB) or \
            os.path.getmtime(_LIB) < max(os.path.getmtime(src),
                                         os.path.getmtime(os.path.join(_HERE, "clo_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB


class _SortSpec(ctypes.Structure):
    _fields_ = [("elem_type", ctypes.c_int), ("key_type", ctypes.c_int),
                ("shift", ctypes.c_uint32), ("mask", ctypes.c_uint64),
                ("descending", ctypes.c_int)]


class _Grand(ctypes.Structure):
    _fields_ = [("mt", ctypes.c_uint32 * 624), ("mti", ctypes.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        vp, sz, u64, u32, i = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64,
                               ctypes.c_uint32, ctypes.c_int)
        L.orc_grand_seed.argtypes = [ctypes.POINTER(_Grand), u32]
        L.orc_grand_int.argtypes = [ctypes.POINTER(_Grand)]
        L.orc_grand_int.restype = u32
        L.orc_grand_double.argtypes = [ctypes.POINTER(_Grand)]
        L.orc_grand_double.restype = ctypes.c_double
        L.orc_grand_int_range.argtypes = [ctypes.POINTER(_Grand), ctypes.c_int32, ctypes.c_int32]
        L.orc_grand_int_range.restype = ctypes.c_int32
        L.orc_fill_sort_input.argtypes = [u32, i, vp, sz]
        L.orc_fill_scan_input.argtypes = [u32, i, vp, sz]
        L.orc_rng_seed_size.argtypes = [i]
        L.orc_rng_seed_size.restype = sz
        L.orc_rng_seed_dev_gid.argtypes = [i, i, u64, u64, sz, vp]
        L.orc_rng_seed_host_mt.argtypes = [i, u64, sz, vp]
        L.orc_rng_generate.argtypes = [i, vp, sz, sz, u32, u32, vp]
        L.orc_scan.argtypes = [i, i, vp, vp, sz]
        L.orc_scan_f64ref.argtypes = [i, vp, vp, sz]
        L.orc_scan_blelloch_port.argtypes = [i, vp, vp, sz, sz, i]
        L.orc_sort_bitonic.argtypes = [ctypes.POINTER(_SortSpec), vp, sz]
        L.orc_sort_gselect.argtypes = [ctypes.POINTER(_SortSpec), vp, vp, sz]
        L.orc_sort_satradix.argtypes = [ctypes.POINTER(_SortSpec), vp, sz, u32, sz, i]
        L.orc_sort_pairs.argtypes = [i, vp, vp, sz]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- GRand -----------------------------------------------------------------

class GRand:
    """GLib GRand (MT19937) restatement."""

    def __init__(self, seed):
        self._s = _Grand()
        lib().orc_grand_seed(ctypes.byref(self._s), seed & 0xFFFFFFFF)

    def int(self):
        return lib().orc_grand_int(ctypes.byref(self._s))

    def double(self):
        return lib().orc_grand_double(ctypes.byref(self._s))

    def int_range(self, b, e):
        return lib().orc_grand_int_range(ctypes.byref(self._s), b, e)


def sort_input(seed, ctype, n):
    """Input vector exactly as clo_sort_bench fills it (clo_sort_bench.c:190-193)."""
    a = np.empty(n, dtype=NP_TYPES[ctype])
    lib().orc_fill_sort_input(seed, ctype, _ptr(a), n)
    return a


def scan_input(seed, ctype, n):
    """Input vector exactly as clo_scan_bench fills it (clo_scan_bench.c:219-224)."""
    a = np.empty(n, dtype=NP_TYPES[ctype])
    lib().orc_fill_scan_input(seed, ctype, _ptr(a), n)
    return a


# ---- RNG ---------------------------------------------------------------------

def rng_seeds_dev_gid(rng, hash_id, main_seed, count, gid0=0):
    r = RNG_NAMES.index(rng)
    st = np.empty(count * RNG_SEED_SIZE[rng], dtype=np.uint8)
    lib().orc_rng_seed_dev_gid(r, hash_id, main_seed, gid0, count, _ptr(st))
    return st


def rng_seeds_host_mt(rng, main_seed, count):
    r = RNG_NAMES.index(rng)
    st = np.empty(count * RNG_SEED_SIZE[rng], dtype=np.uint8)
    lib().orc_rng_seed_host_mt(r, main_seed, count, _ptr(st))
    return st


def rng_generate(rng, states, count, runs, bits=32, maxint=0):
    """Returns (out[runs, count] uint32, updated states). `states` is not modified."""
    r = RNG_NAMES.index(rng)
    st = np.array(states, dtype=np.uint8, copy=True)
    out = np.empty((runs, count), dtype=np.uint32)
    lib().orc_rng_generate(r, _ptr(st), count, runs, bits, maxint, _ptr(out))
    return out, st


# ---- scan ----------------------------------------------------------------------

def scan(a, elem_type, sum_type):
    a = np.ascontiguousarray(a, dtype=NP_TYPES[elem_type])
    out = np.empty(a.size, dtype=NP_TYPES[sum_type])
    rc = lib().orc_scan(elem_type, sum_type, _ptr(a), _ptr(out), a.size)
    if rc:
        raise ValueError("oracle: unsupported scan types")
    return out


def scan_f64ref(a, elem_type):
    a = np.ascontiguousarray(a, dtype=NP_TYPES[elem_type])
    out = np.empty(a.size, dtype=np.float64)
    lib().orc_scan_f64ref(elem_type, _ptr(a), _ptr(out), a.size)
    return out


def scan_blelloch_port(a, elem_type, lws=256, threads=1):
    a = np.ascontiguousarray(a, dtype=NP_TYPES[elem_type])
    out = np.empty_like(a)
    rc = lib().orc_scan_blelloch_port(elem_type, _ptr(a), _ptr(out), a.size, lws, threads)
    if rc:
        raise ValueError("oracle: blelloch port needs power-of-two n >= 2*lws, uint/float")
    return out


# ---- sort ----------------------------------------------------------------------

def _spec(elem_type, key_type=None, shift=0, mask=None, descending=False):
    return _SortSpec(elem_type, elem_type if key_type is None else key_type, shift,
                     0xFFFFFFFFFFFFFFFF if mask is None else mask, int(descending))


def sort_bitonic(a, elem_type, **kw):
    out = np.array(a, dtype=NP_TYPES[elem_type], copy=True)
    s = _spec(elem_type, **kw)
    if lib().orc_sort_bitonic(ctypes.byref(s), _ptr(out), out.size):
        raise ValueError("oracle: bitonic failed")
    return out


def sort_gselect(a, elem_type, **kw):
    a = np.ascontiguousarray(a, dtype=NP_TYPES[elem_type])
    out = np.empty_like(a)
    s = _spec(elem_type, **kw)
    if lib().orc_sort_gselect(ctypes.byref(s), _ptr(a), _ptr(out), a.size):
        raise ValueError("oracle: gselect failed")
    return out


def sort_satradix(a, elem_type, radix=16, lws=256, threads=1, **kw):
    out = np.array(a, dtype=NP_TYPES[elem_type], copy=True)
    s = _spec(elem_type, **kw)
    if lib().orc_sort_satradix(ctypes.byref(s), _ptr(out), out.size, radix, lws, threads):
        raise ValueError("oracle: satradix failed")
    return out


def sort_pairs(keys, payload, key_type):
    k = np.array(keys, dtype=NP_TYPES[key_type], copy=True)
    p = np.array(payload, dtype=np.uint32, copy=True)
    if lib().orc_sort_pairs(key_type, _ptr(k), _ptr(p), k.size):
        raise ValueError("oracle: pairs failed")
    return k, p
