"""cl_ops_b200 -- Python host mirror of the cl_ops C API (B200 backend).

Everything here is a thin ctypes layer over ``libcl_ops.so`` (built from
``cl_ops_b200/csrc`` for sm_100a).  The classes mirror the reference's objects
and entry points one to one (reference: /root/reference/src/cl_ops/):

=================  ==========================================================
``CloSort``        clo_sort_new / clo_sort_with_host_data /
                   clo_sort_with_device_data      (sort/clo_sort_abstract.c:91-418)
``CloScan``        clo_scan_new / clo_scan_with_host_data /
                   clo_scan_with_device_data      (scan/clo_scan_abstract.c:74-362)
``CloRng``         clo_rng_new / clo_rng_get_device_seeds / clo_rng_get_size
                                                  (rng/clo_rng.c:262-481)
``Context`` ...    the cf4ocl2 handles the API takes, mapped onto CUDA
=================  ==========================================================

There is no CPU fallback: importing this package without the built library, or
calling it without a CUDA device, raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcl_ops.so")

# CloType (include/cl_ops/clo_common.h)
CHAR, UCHAR, SHORT, USHORT, INT, UINT, LONG, ULONG, HALF, FLOAT, DOUBLE = range(11)
TYPE_NAMES = ["char", "uchar", "short", "ushort", "int", "uint", "long", "ulong",
              "half", "float", "double"]
NP_TYPES = {CHAR: np.int8, UCHAR: np.uint8, SHORT: np.int16, USHORT: np.uint16,
            INT: np.int32, UINT: np.uint32, LONG: np.int64, ULONG: np.uint64,
            FLOAT: np.float32, DOUBLE: np.float64}
TYPE_SIZES = [1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8]

# clo_error_codes
CLO_SUCCESS, CLO_ERROR_OPENFILE, CLO_ERROR_ARGS, CLO_ERROR_STREAM_WRITE = 0, 1, 2, 3
CLO_ERROR_IMPL_NOT_FOUND, CLO_ERROR_UNKNOWN_TYPE, CLO_ERROR_LIBRARY = 5, 6, 7

# CloRngSeedType
SEED_DEV_GID, SEED_HOST_MT, SEED_EXT_DEV, SEED_EXT_HOST = 0, 1, 2, 3

CL_QUEUE_PROFILING_ENABLE = 1 << 1
CL_MEM_READ_WRITE = 1 << 0


class CloError(RuntimeError):
    """A GError raised by the library (domain CLO_ERROR)."""

    def __init__(self, code, message):
        super().__init__("cl_ops error %d: %s" % (code, message))
        self.code = code
        self.message = message


class _GError(ctypes.Structure):
    _fields_ = [("domain", ctypes.c_uint32), ("code", ctypes.c_int),
                ("message", ctypes.c_char_p)]


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "cl_ops_b200: %s is missing -- build it with `make -C cl_ops_b200/csrc` "
                "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, u64, u32, i = (ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64,
                               ctypes.c_uint32, ctypes.c_int)
        cp = ctypes.c_char_p
        errp = ctypes.POINTER(ctypes.POINTER(_GError))

        def sig(name, restype, *argtypes):
            f = getattr(L, name)
            f.restype = restype
            f.argtypes = list(argtypes)

        sig("clo_b200_version", cp)
        sig("clo_b200_launch_count", u64)
        sig("clo_b200_error_free", None, ctypes.POINTER(_GError))
        sig("clo_type_get_name", cp, i)
        sig("clo_type_sizeof", sz, i)
        sig("clo_type_by_name", i, cp, errp)
        sig("clo_nlpo2", u32, u32)
        sig("clo_ones32", u32, u32)
        sig("clo_tzc", u32, i)
        sig("clo_sum", u32, u32)
        sig("ccl_context_new_from_menu_full", vp, vp, errp)
        sig("ccl_context_new_any", vp, errp)
        sig("ccl_context_destroy", None, vp)
        sig("ccl_queue_new", vp, vp, vp, u64, errp)
        sig("ccl_queue_new_wrap", vp, vp, vp, errp)
        sig("ccl_queue_destroy", None, vp)
        sig("ccl_queue_finish", u32, vp, errp)
        sig("ccl_queue_gc", None, vp)
        sig("ccl_buffer_new", vp, vp, u64, sz, vp, errp)
        sig("ccl_buffer_new_wrap", vp, vp, vp, sz, errp)
        sig("ccl_buffer_destroy", None, vp)
        sig("ccl_buffer_get_ptr", vp, vp)
        sig("ccl_buffer_enqueue_write", vp, vp, vp, u32, sz, sz, vp, vp, errp)
        sig("ccl_buffer_enqueue_read", vp, vp, vp, u32, sz, sz, vp, vp, errp)
        sig("ccl_event_get_duration_ns", u64, vp, errp)
        sig("ccl_wrapper_memcheck", u32)
        sig("clo_sort_new", vp, cp, cp, vp, ctypes.POINTER(i), ctypes.POINTER(i), cp, cp, cp, errp)
        sig("clo_sort_destroy", None, vp)
        sig("clo_sort_with_device_data", vp, vp, vp, vp, vp, vp, sz, sz, errp)
        sig("clo_sort_with_host_data", u32, vp, vp, vp, vp, vp, sz, sz, errp)
        sig("clo_sort_pairs_with_device_data", vp, vp, vp, vp, vp, sz, errp)
        sig("clo_sort_partition_with_device_data", vp, vp, vp, vp, vp, vp, vp, sz, u64, vp, vp, u32, vp, errp)
        sig("clo_sort_partition_count_with_device_data", vp, vp, vp, vp, sz, u64, vp, vp, u32, vp, errp)
        sig("clo_sort_partition_scatter_with_device_data", vp, vp, vp, vp, vp, sz, u64, vp, vp, u32, vp, vp, vp, vp, errp)
        sig("clo_b200_ipc_export", u32, vp, ctypes.c_char_p, errp)
        sig("clo_b200_ipc_import", vp, vp, ctypes.c_char_p, sz, errp)
        sig("clo_sort_b200_debug", u32, vp, vp, ctypes.POINTER(ctypes.c_uint64))
        sig("clo_sort_b200_set_timing", None, vp, u32)
        sig("clo_sort_b200_get_timing", u32, vp, ctypes.POINTER(ctypes.c_float), u32)
        sig("clo_sort_get_element_size", sz, vp)
        sig("clo_sort_get_key_size", sz, vp)
        sig("clo_sort_get_num_kernels", u32, vp, errp)
        sig("clo_sort_get_kernel_name", cp, vp, u32, errp)
        sig("clo_sort_get_localmem_usage", sz, vp, u32, sz, sz, errp)
        sig("clo_scan_new", vp, cp, cp, vp, i, i, cp, errp)
        sig("clo_scan_destroy", None, vp)
        sig("clo_scan_with_device_data", vp, vp, vp, vp, vp, vp, sz, sz, errp)
        sig("clo_scan_with_host_data", u32, vp, vp, vp, vp, vp, sz, sz, errp)
        sig("clo_scan_reduce_with_device_data", vp, vp, vp, vp, vp, sz, errp)
        sig("clo_scan_with_device_data_carry", vp, vp, vp, vp, vp, vp, sz, errp)
        sig("clo_scan_get_num_kernels", u32, vp, errp)
        sig("clo_scan_get_kernel_name", cp, vp, u32, errp)
        sig("clo_rng_new", vp, cp, i, vp, sz, u64, cp, vp, vp, errp)
        sig("clo_rng_new_dev_gid_offset", vp, cp, sz, u64, u64, cp, vp, vp, errp)
        sig("clo_rng_destroy", None, vp)
        sig("clo_rng_get_source", cp, vp)
        sig("clo_rng_get_device_seeds", vp, vp)
        sig("clo_rng_get_size", sz, vp)
        sig("clo_rng_generate", vp, vp, vp, vp, sz, u32, u32, errp)
        sig("clo_rng_generate_host", u32, vp, vp, vp, sz, u32, u32, errp)
        sig("clo_dist_new", vp, vp, vp, errp)
        sig("clo_dist_destroy", None, vp)
        sig("clo_dist_sort_setup", u32, vp, i, sz, u32, errp)
        sig("clo_dist_sort_with_device_data", u32, vp, vp, vp, vp, sz, u64, vp, vp, sz, ctypes.POINTER(sz), errp)
        sig("clo_dist_scan_with_device_data", vp, vp, vp, vp, vp, vp, sz, errp)
        sig("clo_dist_rng_partition", None, u64, u32, u32, ctypes.POINTER(u64), ctypes.POINTER(u64))
        sig("clo_dist_set_timing", None, vp, u32)
        sig("clo_dist_get_phases", u32, vp, ctypes.POINTER(ctypes.c_float), u32)
        sig("clo_dist_get_counts", u32, vp, ctypes.POINTER(u64), ctypes.POINTER(u64))
        _lib = L
    return _lib


class _Err:
    """GError** out-parameter; raises CloError when the call set it."""

    def __init__(self):
        self.p = ctypes.POINTER(_GError)()

    def ref(self):
        return ctypes.byref(self.p)

    def check(self):
        if self.p:
            code = self.p.contents.code
            msg = (self.p.contents.message or b"").decode("utf-8", "replace")
            lib().clo_b200_error_free(self.p)
            self.p = ctypes.POINTER(_GError)()
            raise CloError(code, msg)


def _b(s):
    return None if s is None else s.encode()


def _np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------
# cf4ocl2-style handles
# --------------------------------------------------------------------------

class Context:
    """CCLContext: one CUDA device (ccl_context_new_from_menu_full)."""

    def __init__(self, device_index=None):
        e = _Err()
        if device_index is None:
            self.h = lib().ccl_context_new_any(e.ref())
        else:
            idx = ctypes.c_int(device_index)
            self.h = lib().ccl_context_new_from_menu_full(ctypes.byref(idx), e.ref())
        e.check()

    def destroy(self):
        if self.h:
            lib().ccl_context_destroy(self.h)
            self.h = None


class Queue:
    """CCLQueue: a CUDA stream (ccl_queue_new) or a wrapped existing stream."""

    def __init__(self, ctx, profiling=False, stream=None):
        e = _Err()
        self.ctx = ctx
        if stream is None:
            self.h = lib().ccl_queue_new(ctx.h, None, CL_QUEUE_PROFILING_ENABLE if profiling else 0, e.ref())
        else:
            self.h = lib().ccl_queue_new_wrap(ctx.h, ctypes.c_void_p(stream), e.ref())
        e.check()

    def finish(self):
        e = _Err()
        lib().ccl_queue_finish(self.h, e.ref())
        e.check()

    def gc(self):
        lib().ccl_queue_gc(self.h)

    def destroy(self):
        if self.h:
            lib().ccl_queue_destroy(self.h)
            self.h = None


class Buffer:
    """CCLBuffer: device memory owned by the library, or a wrapped device pointer."""

    def __init__(self, ctx, size=None, ptr=None, keepalive=None):
        e = _Err()
        self.ctx = ctx
        self.size = size
        self._keep = keepalive
        if ptr is None:
            self.h = lib().ccl_buffer_new(ctx.h, CL_MEM_READ_WRITE, size, None, e.ref())
        else:
            self.h = lib().ccl_buffer_new_wrap(ctx.h, ctypes.c_void_p(ptr), size, e.ref())
        e.check()

    @classmethod
    def wrap_tensor(cls, ctx, t):
        """Wrap a contiguous CUDA torch tensor (no copy; the tensor is kept alive)."""
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("wrap_tensor needs a contiguous CUDA tensor")
        return cls(ctx, size=t.numel() * t.element_size(), ptr=t.data_ptr(), keepalive=t)

    @property
    def ptr(self):
        return lib().ccl_buffer_get_ptr(self.h)

    def write(self, queue, host_array, blocking=True):
        a = np.ascontiguousarray(host_array)
        e = _Err()
        lib().ccl_buffer_enqueue_write(self.h, queue.h, 1 if blocking else 0, 0, a.nbytes, _np_ptr(a), None, e.ref())
        e.check()

    def read(self, queue, dtype, count):
        out = np.empty(count, dtype=dtype)
        e = _Err()
        lib().ccl_buffer_enqueue_read(self.h, queue.h, 1, 0, out.nbytes, _np_ptr(out), None, e.ref())
        e.check()
        return out

    def ipc_export(self):
        """64-byte CUDA IPC handle of a library-owned buffer (for a peer process on this box)."""
        out = ctypes.create_string_buffer(64)
        e = _Err()
        lib().clo_b200_ipc_export(self.h, out, e.ref())
        e.check()
        return out.raw

    @classmethod
    def ipc_import(cls, ctx, handle, size):
        """Map a peer process's exported buffer; destroy() unmaps it."""
        b = cls.__new__(cls)
        b.ctx, b.size, b._keep = ctx, size, None
        e = _Err()
        b.h = lib().clo_b200_ipc_import(ctx.h, ctypes.c_char_p(bytes(handle)), size, e.ref())
        e.check()
        return b

    def destroy(self):
        if self.h:
            lib().ccl_buffer_destroy(self.h)
            self.h = None
            self._keep = None


def event_duration_ns(evt):
    e = _Err()
    ns = lib().ccl_event_get_duration_ns(evt, e.ref())
    e.check()
    return ns


# --------------------------------------------------------------------------
# CloSort
# --------------------------------------------------------------------------

class CloSort:
    """Sorter object (clo_sort_new).  `type` in sbitonic|abitonic|gselect|satradix."""

    def __init__(self, type, ctx, elem_type, key_type=None, options=None, compare=None,
                 get_key=None, compiler_opts=None):
        e = _Err()
        et = ctypes.c_int(elem_type)
        kt = ctypes.c_int(key_type) if key_type is not None else None
        self.ctx = ctx
        self.elem_type = elem_type
        self.h = lib().clo_sort_new(_b(type), _b(options), ctx.h, ctypes.byref(et),
                                    ctypes.byref(kt) if kt is not None else None,
                                    _b(compare), _b(get_key), _b(compiler_opts), e.ref())
        e.check()

    def with_host_data(self, data, queue=None, lws_max=0):
        """clo_sort_with_host_data: returns the sorted copy of `data` (numpy)."""
        a = np.ascontiguousarray(data, dtype=NP_TYPES[self.elem_type])
        out = np.empty_like(a)
        e = _Err()
        ok = lib().clo_sort_with_host_data(self.h, queue.h if queue else None, None, _np_ptr(a),
                                           _np_ptr(out), a.size, lws_max, e.ref())
        e.check()
        if not ok:
            raise CloError(CLO_ERROR_LIBRARY, "clo_sort_with_host_data failed")
        return out

    def with_host_pointers(self, in_ptr, out_ptr, numel, queue=None, lws_max=0):
        """clo_sort_with_host_data on raw host addresses (e.g. pinned torch tensors)."""
        e = _Err()
        ok = lib().clo_sort_with_host_data(self.h, queue.h if queue else None, None, ctypes.c_void_p(in_ptr),
                                           ctypes.c_void_p(out_ptr), numel, lws_max, e.ref())
        e.check()
        if not ok:
            raise CloError(CLO_ERROR_LIBRARY, "clo_sort_with_host_data failed")

    def with_device_data(self, queue, data_in, data_out, numel, lws_max=0):
        """clo_sort_with_device_data: data_out=None sorts in place. Returns the CCLEvent*."""
        e = _Err()
        evt = lib().clo_sort_with_device_data(self.h, queue.h, None, data_in.h,
                                              data_out.h if data_out else None, numel, lws_max, e.ref())
        e.check()
        return evt

    def pairs_with_device_data(self, queue, keys, payload, numel):
        e = _Err()
        evt = lib().clo_sort_pairs_with_device_data(self.h, queue.h, keys.h, payload.h, numel, e.ref())
        e.check()
        return evt

    def partition_with_device_data(self, queue, keys_in, payload_in, keys_out, payload_out, numel,
                                   gidx0, splitter_keys, splitter_idx, nparts, counts_out):
        e = _Err()
        evt = lib().clo_sort_partition_with_device_data(
            self.h, queue.h, keys_in.h, payload_in.h if payload_in else None, keys_out.h,
            payload_out.h if payload_out else None, numel, gidx0,
            splitter_keys.h if splitter_keys else None, splitter_idx.h if splitter_idx else None,
            nparts, counts_out.h, e.ref())
        e.check()
        return evt

    def partition_count_with_device_data(self, queue, keys_in, numel, gidx0, splitter_keys, splitter_idx,
                                         nparts, counts_out):
        e = _Err()
        evt = lib().clo_sort_partition_count_with_device_data(
            self.h, queue.h, keys_in.h, numel, gidx0, splitter_keys.h if splitter_keys else None,
            splitter_idx.h if splitter_idx else None, nparts, counts_out.h, e.ref())
        e.check()
        return evt

    def partition_scatter_with_device_data(self, queue, keys_in, payload_in, numel, gidx0, splitter_keys,
                                           splitter_idx, nparts, first_slot, dest_ptrs, payload_dest_ptrs, ok_flag):
        e = _Err()
        evt = lib().clo_sort_partition_scatter_with_device_data(
            self.h, queue.h, keys_in.h, payload_in.h if payload_in else None, numel, gidx0,
            splitter_keys.h if splitter_keys else None, splitter_idx.h if splitter_idx else None, nparts,
            first_slot.h, dest_ptrs.h, payload_dest_ptrs.h if payload_dest_ptrs else None,
            ok_flag.h if ok_flag else None, e.ref())
        e.check()
        return evt

    def set_timing(self, on=True):
        lib().clo_sort_b200_set_timing(self.h, 1 if on else 0)

    def get_timing(self):
        """[histogram+scan ms, pass0 ms, pass1 ms, ...] of the last radix call."""
        out = (ctypes.c_float * 12)()
        k = lib().clo_sort_b200_get_timing(self.h, out, 12)
        return [float(out[i]) for i in range(k)]

    def debug(self, queue):
        """[timeout flag, repaired tiles, 16 phase-profile counters] of the last radix call."""
        out = (ctypes.c_uint64 * 18)()
        lib().clo_sort_b200_debug(self.h, queue.h, out)
        return list(out)

    def kernel_names(self):
        e = _Err()
        n = lib().clo_sort_get_num_kernels(self.h, e.ref())
        return [lib().clo_sort_get_kernel_name(self.h, k, e.ref()).decode() for k in range(n)]

    def destroy(self):
        if self.h:
            lib().clo_sort_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------
# CloScan
# --------------------------------------------------------------------------

class CloScan:
    """Scanner object (clo_scan_new).  `type` is "blelloch"."""

    def __init__(self, type, ctx, elem_type, sum_type, options=None, compiler_opts=None):
        e = _Err()
        self.ctx = ctx
        self.elem_type, self.sum_type = elem_type, sum_type
        self.h = lib().clo_scan_new(_b(type), _b(options), ctx.h, elem_type, sum_type,
                                    _b(compiler_opts), e.ref())
        e.check()

    def with_host_data(self, data, queue=None, lws_max=0):
        a = np.ascontiguousarray(data, dtype=NP_TYPES[self.elem_type])
        out = np.empty(a.size, dtype=NP_TYPES[self.sum_type])
        e = _Err()
        ok = lib().clo_scan_with_host_data(self.h, queue.h if queue else None, None, _np_ptr(a),
                                           _np_ptr(out), a.size, lws_max, e.ref())
        e.check()
        if not ok:
            raise CloError(CLO_ERROR_LIBRARY, "clo_scan_with_host_data failed")
        return out

    def with_host_pointers(self, in_ptr, out_ptr, numel, queue=None, lws_max=0):
        """clo_scan_with_host_data on raw host addresses (e.g. pinned torch tensors)."""
        e = _Err()
        ok = lib().clo_scan_with_host_data(self.h, queue.h if queue else None, None, ctypes.c_void_p(in_ptr),
                                           ctypes.c_void_p(out_ptr), numel, lws_max, e.ref())
        e.check()
        if not ok:
            raise CloError(CLO_ERROR_LIBRARY, "clo_scan_with_host_data failed")

    def with_device_data(self, queue, data_in, data_out, numel, lws_max=0, carry_in=None):
        e = _Err()
        if carry_in is None:
            evt = lib().clo_scan_with_device_data(self.h, queue.h, None, data_in.h, data_out.h,
                                                  numel, lws_max, e.ref())
        else:
            evt = lib().clo_scan_with_device_data_carry(self.h, queue.h, data_in.h, data_out.h,
                                                        carry_in.h, numel, e.ref())
        e.check()
        return evt

    def reduce_with_device_data(self, queue, data_in, total_out, numel):
        e = _Err()
        evt = lib().clo_scan_reduce_with_device_data(self.h, queue.h, data_in.h, total_out.h, numel, e.ref())
        e.check()
        return evt

    def destroy(self):
        if self.h:
            lib().clo_scan_destroy(self.h)
            self.h = None


# --------------------------------------------------------------------------
# CloRng
# --------------------------------------------------------------------------

RNG_SEED_SIZE = {"lcg": 8, "xorshift64": 8, "xorshift128": 16, "mwc64x": 8,
                 "parkmiller": 4, "tauslcg": 16}


class CloRng:
    """RNG object (clo_rng_new): device seeds + generator."""

    def __init__(self, type, ctx, seed_type=SEED_DEV_GID, seeds=None, seeds_count=0, main_seed=0,
                 hash=None, queue=None, gid_offset=None):
        e = _Err()
        self.ctx = ctx
        self.type = type
        self.seeds_count = seeds_count
        self._keep = None
        if gid_offset is not None:
            self.h = lib().clo_rng_new_dev_gid_offset(_b(type), seeds_count, gid_offset, main_seed,
                                                      _b(hash), ctx.h, queue.h if queue else None, e.ref())
        else:
            sp = None
            if seed_type == SEED_EXT_HOST and seeds is not None:
                self._keep = np.ascontiguousarray(seeds)
                sp = _np_ptr(self._keep)
            elif seed_type == SEED_EXT_DEV and seeds is not None:
                self._keep = seeds
                sp = seeds.h
            self.h = lib().clo_rng_new(_b(type), seed_type, sp, seeds_count, main_seed, _b(hash),
                                       ctx.h, queue.h if queue else None, e.ref())
        e.check()

    def get_source(self):
        return lib().clo_rng_get_source(self.h).decode()

    def get_size(self):
        return lib().clo_rng_get_size(self.h)

    def get_device_seeds_ptr(self):
        return lib().ccl_buffer_get_ptr(lib().clo_rng_get_device_seeds(self.h))

    def read_seeds(self, queue):
        """Copy the device states back (raw bytes)."""
        out = np.empty(self.get_size(), dtype=np.uint8)
        e = _Err()
        lib().ccl_buffer_enqueue_read(lib().clo_rng_get_device_seeds(self.h), queue.h, 1, 0,
                                      out.nbytes, _np_ptr(out), None, e.ref())
        e.check()
        return out

    def generate(self, queue, out, runs, bits=32, maxint=0):
        """Bulk generation into a device Buffer: out[r*G+g] (clo_rng_generate)."""
        e = _Err()
        evt = lib().clo_rng_generate(self.h, queue.h, out.h, runs, bits, maxint, e.ref())
        e.check()
        return evt

    def generate_host(self, runs, bits=32, maxint=0, queue=None):
        out = np.empty((runs, self.seeds_count), dtype=np.uint32)
        e = _Err()
        ok = lib().clo_rng_generate_host(self.h, queue.h if queue else None, _np_ptr(out), runs,
                                         bits, maxint, e.ref())
        e.check()
        if not ok:
            raise CloError(CLO_ERROR_LIBRARY, "clo_rng_generate_host failed")
        return out

    def destroy(self):
        if self.h:
            lib().clo_rng_destroy(self.h)
            self.h = None


def launch_count():
    return lib().clo_b200_launch_count()


def version():
    return lib().clo_b200_version().decode()


# --------------------------------------------------------------------------
# CloDist: the multi-GPU entry points (include/cl_ops/clo_b200.h, csrc/dist.cu)
# --------------------------------------------------------------------------

_AG_DEV = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
_BAR_DEV = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p)
_AG_HOST = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)


class _DistComm(ctypes.Structure):
    _fields_ = [("user", ctypes.c_void_p), ("rank", ctypes.c_uint32), ("world", ctypes.c_uint32),
                ("all_gather_dev", _AG_DEV), ("barrier_dev", _BAR_DEV), ("all_gather_host", _AG_HOST)]


class _RawDev:
    """device memory at a raw address as a torch tensor (zero copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class CloDist:
    """clo_dist_*: sample sort / scan / RNG partition over the GPUs of one box, one process per
    GPU.  The three collectives the library asks for are served by torch.distributed (NCCL) on
    torch's CURRENT stream, so `queue` must wrap that stream (Queue(ctx, stream=...))."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self.group, self.ctx = torch, dist, group, ctx
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._views = {}
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda")

        def view(ptr, nbytes):
            key = (ptr, nbytes)
            t = self._views.get(key)
            if t is None:
                t = torch.as_tensor(_RawDev(ptr, nbytes), device=torch.device("cuda", torch.cuda.current_device()))
                self._views[key] = t
            return t

        def ag_dev(user, send, recv, nbytes, stream):
            try:
                dist.all_gather_into_tensor(view(recv, nbytes * self.world), view(send, nbytes), group=group)
                return 0
            except Exception:          # the C side turns this into a GError
                return 1

        def bar_dev(user, stream):
            try:
                dist.all_reduce(self._flag, group=group)
                return 0
            except Exception:
                return 1

        def ag_host(user, send, recv, nbytes):
            try:
                mine = ctypes.string_at(send, nbytes)
                parts = [None] * self.world
                dist.all_gather_object(parts, mine, group=group)
                ctypes.memmove(recv, b"".join(parts), nbytes * self.world)
                return 0
            except Exception:
                return 1

        self._cbs = (_AG_DEV(ag_dev), _BAR_DEV(bar_dev), _AG_HOST(ag_host))      # keep the thunks alive
        self._comm = _DistComm(None, self.rank, self.world, *self._cbs)
        e = _Err()
        self.h = lib().clo_dist_new(ctx.h, ctypes.byref(self._comm), e.ref())
        e.check()

    def sort_setup(self, key_type, capacity, with_payload=False):
        e = _Err()
        lib().clo_dist_sort_setup(self.h, key_type, capacity, 1 if with_payload else 0, e.ref())
        e.check()

    def sort(self, queue, keys_in, payload_in, numel, keys_out, payload_out, out_capacity, gidx0=None):
        """clo_dist_sort_with_device_data on Buffers; returns the number of elements received."""
        n_out = ctypes.c_size_t(0)
        e = _Err()
        lib().clo_dist_sort_with_device_data(self.h, queue.h, keys_in.h, payload_in.h if payload_in else None, numel,
                                             0xFFFFFFFFFFFFFFFF if gidx0 is None else gidx0, keys_out.h,
                                             payload_out.h if payload_out else None, out_capacity, ctypes.byref(n_out), e.ref())
        e.check()
        return n_out.value

    def scan(self, scanner, queue, buf_in, buf_out, numel):
        e = _Err()
        lib().clo_dist_scan_with_device_data(self.h, scanner.h, queue.h, buf_in.h, buf_out.h, numel, e.ref())
        e.check()

    @staticmethod
    def rng_partition(total_streams, rank, world):
        first, count = ctypes.c_uint64(0), ctypes.c_uint64(0)
        lib().clo_dist_rng_partition(total_streams, rank, world, ctypes.byref(first), ctypes.byref(count))
        return first.value, count.value

    def set_timing(self, on):
        lib().clo_dist_set_timing(self.h, 1 if on else 0)

    PHASES = ("samples+allgather", "splitters", "count", "sizes all-gather", "scatter to peers",
              "barrier+sizes to host", "local sort")

    def phases_ms(self):
        out = (ctypes.c_float * 8)()
        n = lib().clo_dist_get_phases(self.h, out, 8)
        return {self.PHASES[k]: float(out[k]) for k in range(n)}

    def counts(self):
        s, r = (ctypes.c_uint64 * 16)(), (ctypes.c_uint64 * 16)()
        lib().clo_dist_get_counts(self.h, s, r)
        return list(s)[:self.world], list(r)[:self.world]

    def destroy(self):
        if self.h:
            lib().clo_dist_destroy(self.h)
            self.h = None
            self._views = {}
