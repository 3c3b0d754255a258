"""GPU parity: clo_scan (single-pass look-back kernel) vs the CPU oracle, through the C-ABI.

Bar: bit-exact for integer sums (any wrap-around included); float sums within
|gpu - ref| <= 1e-5 * |ref| + 1e-3 of a double-precision host prefix sum.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 31, 257, 4096, 4097, 12289, 100003, (1 << 20) + 3, (1 << 22) + 5]
INT_PAIRS = [(oracle.UINT, oracle.UINT), (oracle.UINT, oracle.ULONG), (oracle.INT, oracle.LONG),
             (oracle.UCHAR, oracle.UINT), (oracle.USHORT, oracle.ULONG), (oracle.ULONG, oracle.ULONG),
             (oracle.CHAR, oracle.INT), (oracle.LONG, oracle.UINT), (oracle.SHORT, oracle.USHORT),
             (oracle.UINT, oracle.UCHAR)]


def _rand(rng, ctype, n, full_range):
    dt = oracle.NP_TYPES[ctype]
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        if full_range:
            return rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
        return rng.integers(0, 128, size=n).astype(dt)
    return rng.random(n).astype(dt)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("et,st", INT_PAIRS)
def test_scan_int_bit_exact(clo, ctx, queue, et, st, n):
    rng = np.random.default_rng(1000 + n + 31 * et + 7 * st)
    for full_range in (False, True):
        a = _rand(rng, et, n, full_range)
        s = clo.CloScan("blelloch", ctx, et, st)
        got = s.with_host_data(a, queue)
        s.destroy()
        want = oracle.scan(a, et, st)
        assert got.dtype == want.dtype
        assert np.array_equal(got, want), "scan %s->%s n=%d full=%s first diff at %d" % (
            oracle.TYPE_NAMES[et], oracle.TYPE_NAMES[st], n, full_range,
            int(np.flatnonzero(got != want)[0]))


def test_scan_bench_input_matches_serial_host_scan(clo, ctx, queue):
    """The reference's own acceptance check (clo_scan_bench.c:246-271): uint -> ulong,
    values in [0,128) from GRand seed 0, exact equality with a serial host scan."""
    a = oracle.scan_input(0, oracle.UINT, 1 << 18)
    s = clo.CloScan("blelloch", ctx, oracle.UINT, oracle.ULONG)
    got = s.with_host_data(a, queue)
    s.destroy()
    want = np.concatenate(([0], np.cumsum(a.astype(np.uint64))[:-1])).astype(np.uint64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [1, 1000, 4097, (1 << 20) + 3, 1 << 22])
@pytest.mark.parametrize("et,st", [(oracle.FLOAT, oracle.FLOAT), (oracle.FLOAT, oracle.DOUBLE),
                                   (oracle.DOUBLE, oracle.DOUBLE), (oracle.UINT, oracle.FLOAT)])
def test_scan_float_tolerance(clo, ctx, queue, et, st, n):
    rng = np.random.default_rng(77 + n)
    a = _rand(rng, et, n, False)
    s = clo.CloScan("blelloch", ctx, et, st)
    got = s.with_host_data(a, queue).astype(np.float64)
    s.destroy()
    ref = oracle.scan_f64ref(a, et)
    tol = 1e-5 * np.abs(ref) + 1e-3
    err = np.abs(got - ref)
    assert np.all(err <= tol), "max err %.3e (rel %.3e)" % (err.max(), (err / np.maximum(np.abs(ref), 1)).max())


def test_scan_device_data_carry_and_reduce(clo, ctx, queue):
    """device-data entry point on wrapped torch memory; carry-in and reduce are the
    building blocks of the multi-GPU scan."""
    import torch
    n = (1 << 21) + 17
    a = oracle.scan_input(3, oracle.UINT, n)
    t_in = torch.from_numpy(a.view(np.int32)).cuda()
    t_out = torch.empty(n, dtype=torch.int64, device="cuda")
    t_carry = torch.tensor([123456789012], dtype=torch.int64, device="cuda")
    t_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
    b_carry, b_total = clo.Buffer.wrap_tensor(ctx, t_carry), clo.Buffer.wrap_tensor(ctx, t_total)
    s = clo.CloScan("blelloch", ctx, oracle.UINT, oracle.ULONG)
    s.with_device_data(queue, b_in, b_out, n, carry_in=b_carry)
    s.reduce_with_device_data(queue, b_in, b_total, n)
    queue.finish()
    want = oracle.scan(a, oracle.UINT, oracle.ULONG) + np.uint64(123456789012)
    assert np.array_equal(t_out.cpu().numpy().view(np.uint64), want)
    assert int(t_total.item()) == int(a.astype(np.uint64).sum())
    # second call on the same scanner (epoch / ticket bookkeeping)
    s.with_device_data(queue, b_in, b_out, n)
    queue.finish()
    assert np.array_equal(t_out.cpu().numpy().view(np.uint64), oracle.scan(a, oracle.UINT, oracle.ULONG))
    for b in (b_in, b_out, b_carry, b_total):
        b.destroy()
    s.destroy()


@pytest.mark.parametrize("kernel", ["pp", "classic"])
def test_scan_large_both_kernels(clo, ctx, queue, kernel, monkeypatch):
    """the persistent (propagator) kernel and the one-tile-per-CTA kernel give the same bits;
    carry-in goes through the propagator in the persistent kernel"""
    import torch
    monkeypatch.setenv("CLO_SCAN_KERNEL", kernel)
    n = (1 << 24) + 4099
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    t_in = torch.from_numpy(a.view(np.int32)).cuda()
    t_out = torch.empty(n, dtype=torch.int64, device="cuda")
    t_carry = torch.tensor([2**40 + 7], dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    b_in, b_out, b_c = (clo.Buffer.wrap_tensor(ctx, t) for t in (t_in, t_out, t_carry))
    s = clo.CloScan("blelloch", ctx, oracle.UINT, oracle.ULONG)
    want = oracle.scan(a, oracle.UINT, oracle.ULONG)
    for rep in range(3):      # repeated calls exercise the epoch tags
        s.with_device_data(queue, b_in, b_out, n, carry_in=b_c if rep == 1 else None)
        queue.finish()
        w = want + np.uint64(2**40 + 7) if rep == 1 else want
        assert np.array_equal(t_out.cpu().numpy().view(np.uint64), w), "rep %d" % rep
    for b in (b_in, b_out, b_c):
        b.destroy()
    s.destroy()


@pytest.mark.parametrize("et", [oracle.UINT, oracle.INT, oracle.FLOAT, oracle.ULONG, oracle.LONG, oracle.DOUBLE])
@pytest.mark.parametrize("n", [(1 << 23) + 3 * 4096 + 32, (1 << 22), (1 << 22) + 4096 * 295 + 64])
def test_scan_copy_engine_kernel(clo, ctx, queue, et, n, monkeypatch):
    """clo_scan_tma (element and sum type of the same size, n a multiple of 128 bytes): TMA boxes
    clipped by the tensor map on the ragged last tile, carry-in through the propagator chain,
    repeated calls (epoch tags), in place, and the same bits as the cp.async kernel it replaces."""
    import torch
    dt = oracle.NP_TYPES[et]
    rng = np.random.default_rng(n % 1000 + et)
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        a = rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
    else:
        a = rng.random(n).astype(dt)
    tdt = {np.uint32: torch.int32, np.int32: torch.int32, np.float32: torch.float32, np.uint64: torch.int64,
           np.int64: torch.int64, np.float64: torch.float64}[dt]
    sdt = {4: np.int32, 8: np.int64}[a.dtype.itemsize] if np.issubdtype(dt, np.integer) else dt
    t_in = torch.from_numpy(a.view(sdt).copy()).cuda()
    t_out = torch.empty(n, dtype=tdt, device="cuda")
    carry_val = 12345 if np.issubdtype(dt, np.integer) else 1000.5
    t_carry = torch.tensor([carry_val], dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    b_in, b_out, b_c = (clo.Buffer.wrap_tensor(ctx, t) for t in (t_in, t_out, t_carry))
    s = clo.CloScan("blelloch", ctx, et, et)
    if np.issubdtype(dt, np.integer):
        want = oracle.scan(a, et, et)
        def check(got, carry):
            w = (want.astype(np.uint64) + np.uint64(carry)).astype(want.dtype) if carry else want
            assert np.array_equal(got.view(want.dtype), w)
    else:
        ref = np.cumsum(a.astype(np.float64)) - a.astype(np.float64)
        def check(got, carry):
            r = ref + carry
            tol = (1e-5 if dt == np.float32 else 1e-12) * np.abs(r) + (1e-3 if dt == np.float32 else 1e-9)
            assert np.all(np.abs(got.astype(np.float64) - r) <= tol)
    for rep in range(3):
        s.with_device_data(queue, b_in, b_out, n, carry_in=b_c if rep == 1 else None)
        queue.finish()
        check(t_out.cpu().numpy(), carry_val if rep == 1 else 0)
    first = t_out.cpu().numpy().copy()
    # the cp.async kernel gives the same bits (integers) / the same tolerance (floats)
    monkeypatch.setenv("CLO_SCAN_KERNEL", "pp")
    s2 = clo.CloScan("blelloch", ctx, et, et)
    s2.with_device_data(queue, b_in, b_out, n)
    queue.finish()
    if np.issubdtype(dt, np.integer):
        assert np.array_equal(t_out.cpu().numpy(), first)
    else:
        check(t_out.cpu().numpy(), 0)
    s2.destroy()
    # in place
    s.with_device_data(queue, b_in, b_in, n)
    queue.finish()
    check(t_in.cpu().numpy(), 0)
    for b in (b_in, b_out, b_c):
        b.destroy()
    s.destroy()


@pytest.mark.parametrize("et,st", [(oracle.UINT, oracle.UINT), (oracle.UINT, oracle.ULONG), (oracle.INT, oracle.LONG),
                                   (oracle.FLOAT, oracle.FLOAT), (oracle.UCHAR, oracle.UINT)])
@pytest.mark.parametrize("chunk,n", [("100000", 1234567), ("65536", 65536 * 5), (None, (1 << 25) + 77)])
def test_scan_host_data_is_pipelined_in_chunks(clo, ctx, queue, et, st, chunk, n, monkeypatch):
    """clo_scan_with_host_data on large inputs: chunks go in, are scanned with the carry of everything
    before them and come out while the next chunk is already on its way (both directions of the host
    link busy).  Same bits as the oracle for integers (wrap-around kept across chunk borders)."""
    if chunk is not None:
        monkeypatch.setenv("CLO_SCAN_HOST_CHUNK", chunk)
    rng = np.random.default_rng(n % 977 + et + st)
    dt = oracle.NP_TYPES[et]
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        a = rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
    else:
        a = rng.random(n).astype(dt)
    s = clo.CloScan("blelloch", ctx, et, st)
    for rep in range(2):
        got = s.with_host_data(a, queue)
        if np.issubdtype(dt, np.integer):
            assert np.array_equal(got, oracle.scan(a, et, st))
        else:
            ref = np.cumsum(a.astype(np.float64)) - a
            assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-3)
    # and with the pipeline off: the same result
    monkeypatch.setenv("CLO_SCAN_HOST_CHUNK", "0")
    got0 = s.with_host_data(a, queue)
    if np.issubdtype(dt, np.integer):
        assert np.array_equal(got0, got)
    s.destroy()


def test_scan_errors(clo, ctx):
    with pytest.raises(clo.CloError) as ei:
        clo.CloScan("nosuchscan", ctx, oracle.UINT, oracle.UINT)
    assert ei.value.code == clo.CLO_ERROR_IMPL_NOT_FOUND
    with pytest.raises(clo.CloError) as ei:
        clo.CloScan("blelloch", ctx, oracle.UINT, oracle.UINT, options="foo=1")
    assert ei.value.code == clo.CLO_ERROR_ARGS
