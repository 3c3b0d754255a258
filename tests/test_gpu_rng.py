"""GPU parity: clo_rng seeding and bulk generation vs the CPU oracle -- bit-exact for
every generator, seed mode, hash, and for the states left behind."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

RNGS = oracle.RNG_NAMES
HASHES = [(None, 0), ("KNUTH(x)", 1), ("XS1(x)", 2)]


@pytest.mark.parametrize("rng", RNGS)
@pytest.mark.parametrize("hash_name,hash_id", HASHES)
@pytest.mark.parametrize("G,runs", [(1, 4), (5, 3), (1024, 16), (10000, 7), (65536 + 4, 5)])
def test_dev_gid_streams(clo, ctx, queue, rng, hash_name, hash_id, G, runs):
    main_seed = 1234
    r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, G, main_seed, hash_name, queue)
    st0 = oracle.rng_seeds_dev_gid(rng, hash_id, main_seed, G)
    assert np.array_equal(r.read_seeds(queue), st0), "seed init differs"
    got = r.generate_host(runs, queue=queue)
    want, st1 = oracle.rng_generate(rng, st0, G, runs)
    assert np.array_equal(got, want)
    assert np.array_equal(r.read_seeds(queue), st1), "states after generate differ"
    # a second call continues every stream
    got2 = r.generate_host(2, queue=queue)
    want2, _ = oracle.rng_generate(rng, st1, G, 2)
    assert np.array_equal(got2, want2)
    r.destroy()


@pytest.mark.parametrize("rng", RNGS)
def test_host_mt_and_ext_host_seeds(clo, ctx, queue, rng):
    G = 4096
    r = clo.CloRng(rng, ctx, clo.SEED_HOST_MT, None, G, 0, None, queue)
    st0 = oracle.rng_seeds_host_mt(rng, 0, G)
    assert np.array_equal(r.read_seeds(queue), st0)
    got = r.generate_host(9, queue=queue)
    want, _ = oracle.rng_generate(rng, st0, G, 9)
    assert np.array_equal(got, want)
    r.destroy()
    # EXT_HOST with the byte pattern of the reference's test (test_rng.c:267-268)
    ss = oracle.RNG_SEED_SIZE[rng]
    pat = (((np.arange(G * ss) + 1) * 3) & 0xFF).astype(np.uint8)
    r = clo.CloRng(rng, ctx, clo.SEED_EXT_HOST, pat, G, 0, None, queue)
    got = r.generate_host(5, queue=queue)
    want, _ = oracle.rng_generate(rng, pat, G, 5)
    assert np.array_equal(got, want)
    r.destroy()


@pytest.mark.parametrize("rng", ["lcg", "xorshift128", "mwc64x"])
def test_bits_and_maxint(clo, ctx, queue, rng):
    G = 2048
    st0 = oracle.rng_seeds_dev_gid(rng, 1, 7, G)
    r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, G, 7, "KNUTH(x)", queue)
    got = r.generate_host(6, bits=8, queue=queue)
    want, st1 = oracle.rng_generate(rng, st0, G, 6, bits=8)
    assert np.array_equal(got, want)
    got = r.generate_host(6, maxint=1000, queue=queue)
    want, _ = oracle.rng_generate(rng, st1, G, 6, maxint=1000)
    assert np.array_equal(got, want)
    r.destroy()


def test_gid_offset_partition_equals_whole(clo, ctx, queue):
    """multi-GPU stream partitioning: streams [off, off+G/2) seeded with a gid offset
    reproduce the second half of the single-device streams (no communication)."""
    G = 8192
    whole = clo.CloRng("xorshift128", ctx, clo.SEED_DEV_GID, None, G, 5, "KNUTH(x)", queue)
    w = whole.generate_host(4, queue=queue)
    half = clo.CloRng("xorshift128", ctx, seeds_count=G // 2, main_seed=5, hash="KNUTH(x)", queue=queue,
                      gid_offset=G // 2)
    h = half.generate_host(4, queue=queue)
    assert np.array_equal(w[:, G // 2:], h)
    whole.destroy()
    half.destroy()


def test_rng_errors_and_source(clo, ctx, queue):
    with pytest.raises(clo.CloError) as ei:
        clo.CloRng("nosuchrng", ctx, clo.SEED_DEV_GID, None, 16, 0, None, queue)
    assert ei.value.code == clo.CLO_ERROR_IMPL_NOT_FOUND
    with pytest.raises(clo.CloError) as ei:
        clo.CloRng("lcg", ctx, clo.SEED_EXT_HOST, None, 16, 0, None, queue)
    assert ei.value.code == clo.CLO_ERROR_ARGS
    r = clo.CloRng("mwc64x", ctx, clo.SEED_DEV_GID, None, 16, 0, "(x * 3 + 1)", queue)  # test_rng.c:42 no-op hash
    assert r.get_size() == 16 * 8
    assert "clo_rng_next" in r.get_source()
    assert np.array_equal(r.read_seeds(queue), oracle.rng_seeds_dev_gid("mwc64x", 0, 0, 16))
    r.destroy()


def test_custom_seed_hash_strings_are_compiled_at_run_time(clo, ctx, queue):
    """CLO_RNG_HASH strings outside {none, KNUTH(x), XS1(x)} are built into the seeding kernel at
    run time (clo_rng.c:101-109 does the same with its OpenCL program): checked against the
    formula in 64-bit arithmetic and clo_ulong2statetype of lcg / xorshift128."""
    G, ms = 5000, 99
    gid = np.arange(G, dtype=np.uint64) + np.uint64(ms)
    with np.errstate(over="ignore"):
        want1 = gid * np.uint64(3) + np.uint64(1)
        want2 = ((gid * np.uint64(2654435761)) % np.uint64(1 << 32)) ^ np.uint64(0xABCDEF)
    r = clo.CloRng("lcg", ctx, clo.SEED_DEV_GID, None, G, ms, "x = x * 3 + 1", queue)
    assert np.array_equal(r.read_seeds(queue).view(np.uint64), want1)
    r.destroy()
    r = clo.CloRng("xorshift128", ctx, clo.SEED_DEV_GID, None, G, ms, "KNUTH(x); x ^= 0xABCDEF", queue)
    st = r.read_seeds(queue).view(np.uint32).reshape(G, 4)
    s = want2
    exp = np.stack([s & 0xFFFFFFFF, (s >> np.uint64(16)) & 0xFFFFFFFF, (s >> np.uint64(32)) & 0xFFFFFFFF,
                    (s >> np.uint64(46)) & 0xFFFFFFFF], axis=1).astype(np.uint32)
    assert np.array_equal(st, exp)
    r.destroy()
    with pytest.raises(clo.CloError) as ei:
        clo.CloRng("lcg", ctx, clo.SEED_DEV_GID, None, 16, 0, "x = x +* 2", queue)
    assert ei.value.code == clo.CLO_ERROR_ARGS and "does not compile" in str(ei.value)


@pytest.mark.parametrize("rng", ["lcg", "xorshift128", "mwc64x", "tauslcg"])
def test_next_int_vector_api_through_the_program_shim(clo, ctx, queue, rng):
    """clo_rng_next_int{,2,4,8} (clo_rng_api.cl:33-105): a client kernel appended to
    clo_rng_get_source(), built and launched through the cf4ocl-style program/kernel shim (as
    clo_rng_bench.c:176-312 does).  Work-item g of a width-W call draws from the streams
    g + k * global_size, k < W (clo_rng_workitem.cl:24-32); every lane must match the oracle."""
    import ctypes
    L = clo.lib()
    vp = ctypes.c_void_p
    L.ccl_program_new_from_source.restype = vp
    L.ccl_program_new_from_source.argtypes = [vp, ctypes.c_char_p, vp]
    L.ccl_program_build.restype = ctypes.c_int
    L.ccl_program_build.argtypes = [vp, ctypes.c_char_p, vp]
    L.ccl_program_get_kernel.restype = vp
    L.ccl_program_get_kernel.argtypes = [vp, ctypes.c_char_p, vp]
    L.ccl_program_destroy.argtypes = [vp]
    L.ccl_arg_new.restype = vp
    L.ccl_arg_new.argtypes = [vp, ctypes.c_size_t]
    L.ccl_kernel_enqueue_ndrange.restype = vp
    L.ccl_kernel_enqueue_ndrange.argtypes = [vp, vp, ctypes.c_uint, vp, vp, vp, vp, vp]
    gs, n = 1024, 1000
    r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, 8 * gs, 42, "KNUTH(x)", queue)
    kernels = """
__kernel void k1(__global clo_statetype* st, __global uint* out, uint n) { out[get_global_id(0)] = clo_rng_next_int(st, n); }
__kernel void k2(__global clo_statetype* st, __global uint* out, uint n) {
	uint g = get_global_id(0), gs = get_global_size(0); uint2 v = clo_rng_next_int2(st, n); out[g] = v.x; out[gs + g] = v.y; }
__kernel void k4(__global clo_statetype* st, __global uint* out, uint n) {
	uint g = get_global_id(0), gs = get_global_size(0); uint4 v = clo_rng_next_int4(st, n);
	out[g] = v.x; out[gs + g] = v.y; out[2 * gs + g] = v.z; out[3 * gs + g] = v.w; }
__kernel void k8(__global clo_statetype* st, __global uint* out, uint n) {
	uint g = get_global_id(0), gs = get_global_size(0); uint8 v = clo_rng_next_int8(st, n);
	out[g] = v.s0; out[gs + g] = v.s1; out[2 * gs + g] = v.s2; out[3 * gs + g] = v.s3;
	out[4 * gs + g] = v.s4; out[5 * gs + g] = v.s5; out[6 * gs + g] = v.s6; out[7 * gs + g] = v.s7; }
"""
    err = clo._Err()
    prg = L.ccl_program_new_from_source(ctx.h, (r.get_source() + kernels).encode(), err.ref())
    err.check()
    assert L.ccl_program_build(prg, None, err.ref())
    err.check()
    states = oracle.rng_seeds_dev_gid(rng, 1, 42, 8 * gs)
    seeds_dev = vp(L.clo_rng_get_device_seeds(r.h))
    out = clo.Buffer(ctx, size=8 * gs * 4)
    for name, width in (("k1", 1), ("k2", 2), ("k4", 4), ("k8", 8)):
        k = L.ccl_program_get_kernel(prg, name.encode(), err.ref())
        err.check()
        nn = ctypes.c_uint(n)
        L.ccl_kernel_set_args(vp(k), seeds_dev, vp(out.h), vp(L.ccl_arg_new(ctypes.byref(nn), 4)), vp(None))
        g, l = ctypes.c_size_t(gs), ctypes.c_size_t(128)
        L.ccl_kernel_enqueue_ndrange(k, queue.h, 1, None, ctypes.byref(g), ctypes.byref(l), None, err.ref())
        err.check()
        got = out.read(queue, np.uint32, width * gs)
        want, states = oracle.rng_generate(rng, states, 8 * gs, 1, maxint=n)
        want = np.asarray(want).reshape(-1)
        assert np.array_equal(got, want[: width * gs]), name
        # streams beyond the call's width did not advance: restore them for the next comparison
        st_dev = r.read_seeds(queue)
        assert np.array_equal(st_dev.view(np.uint8).reshape(8 * gs, -1)[: width * gs],
                              np.asarray(states).view(np.uint8).reshape(8 * gs, -1)[: width * gs])
        states = st_dev
    out.destroy()
    L.ccl_program_destroy(prg)
    r.destroy()
