"""Out-of-bounds WRITE check without compute-sanitizer (it is closed on this GPU pool, see
profiles/r02_sanitizer_closed_on_pool.log): every operator writes into the middle of a larger
allocation whose guard zones hold a pattern, and the guards must be intact afterwards.  Sizes
sit on and around the tile / box boundaries of each kernel (the copy-engine scan relies on the
tensor map's extent to clip its last box; the onesweep and the scan rings on ragged last tiles)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GUARD = 4096          # elements on each side
PATTERN = 0x5A


def _guarded(torch, n, dtype):
    t = torch.full((n + 2 * GUARD,), 0, dtype=dtype, device="cuda")
    t.view(torch.uint8).fill_(PATTERN)
    return t, t[GUARD:GUARD + n]


def _intact(torch, t, n):
    b = t.view(torch.uint8)
    es = t.element_size()
    return bool((b[:GUARD * es] == PATTERN).all()) and bool((b[(GUARD + n) * es:] == PATTERN).all())


@pytest.mark.parametrize("n", [(1 << 22), (1 << 22) + 32, (1 << 22) + 4096 + 64, (1 << 22) + 77, 100003, 8192, 1])
@pytest.mark.parametrize("et,st,tdt,odt", [("UINT", "UINT", "int32", "int32"), ("FLOAT", "FLOAT", "float32", "float32"),
                                            ("UINT", "ULONG", "int32", "int64"), ("ULONG", "ULONG", "int64", "int64")])
def test_scan_writes_stay_inside(clo, ctx, queue, n, et, st, tdt, odt):
    import torch
    x = (torch.rand(n, device="cuda") if tdt == "float32" else torch.randint(0, 1000, (n,), device="cuda")).to(getattr(torch, tdt))
    whole, out = _guarded(torch, n, getattr(torch, odt))
    sc = clo.CloScan("blelloch", ctx, getattr(clo, et), getattr(clo, st))
    bi, bo = clo.Buffer.wrap_tensor(ctx, x), clo.Buffer.wrap_tensor(ctx, out)
    torch.cuda.synchronize()                    # the library's queue is its own stream
    sc.with_device_data(queue, bi, bo, n)
    queue.finish()
    assert _intact(torch, whole, n)
    if tdt != "float32":
        ref = torch.cumsum(x.to(torch.int64), 0) - x.to(torch.int64)
        if odt == "int32":
            assert torch.equal(out.to(torch.int64) & 0xFFFFFFFF, ref & 0xFFFFFFFF)
        else:
            assert torch.equal(out, ref)
    bi.destroy(); bo.destroy(); sc.destroy()


@pytest.mark.parametrize("n", [(1 << 20), (1 << 20) + 1, 8192 * 3 + 5, 8191, 1])
@pytest.mark.parametrize("et,tdt", [("UINT", "int32"), ("ULONG", "int64")])
def test_sort_writes_stay_inside(clo, ctx, queue, n, et, tdt):
    import torch
    hi = 2**31 - 1 if tdt == "int32" else 2**62
    x = torch.randint(-hi, hi, (n,), dtype=getattr(torch, tdt), device="cuda")
    whole, out = _guarded(torch, n, getattr(torch, tdt))
    s = clo.CloSort("satradix", ctx, getattr(clo, et))
    bi, bo = clo.Buffer.wrap_tensor(ctx, x), clo.Buffer.wrap_tensor(ctx, out)
    torch.cuda.synchronize()
    s.with_device_data(queue, bi, bo, n)
    queue.finish()
    assert _intact(torch, whole, n)
    a = x.cpu().numpy().view(np.uint32 if tdt == "int32" else np.uint64)
    assert np.array_equal(out.cpu().numpy().view(a.dtype), np.sort(a))
    # pairs, in place, inside guarded allocations
    wk, k = _guarded(torch, n, getattr(torch, tdt))
    wp, p = _guarded(torch, n, torch.int32)
    k.copy_(x); p.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
    bk, bp = clo.Buffer.wrap_tensor(ctx, k), clo.Buffer.wrap_tensor(ctx, p)
    torch.cuda.synchronize()
    s.pairs_with_device_data(queue, bk, bp, n)
    queue.finish()
    assert _intact(torch, wk, n) and _intact(torch, wp, n)
    assert np.array_equal(a[p.cpu().numpy()], k.cpu().numpy().view(a.dtype))
    for b in (bi, bo, bk, bp):
        b.destroy()
    s.destroy()


@pytest.mark.parametrize("name", oracle.RNG_NAMES)
def test_rng_writes_stay_inside(clo, ctx, queue, name):
    import torch
    G, runs = 4096 + 3, 5
    whole, out = _guarded(torch, G * runs, torch.int32)
    r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, G, 11, "KNUTH(x)", queue)
    bo = clo.Buffer.wrap_tensor(ctx, out)
    torch.cuda.synchronize()
    r.generate(queue, bo, runs)
    queue.finish()
    assert _intact(torch, whole, G * runs)
    want, _ = oracle.rng_generate(name, oracle.rng_seeds_dev_gid(name, 1, 11, G), G, runs)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), np.asarray(want).reshape(-1))
    bo.destroy(); r.destroy()
