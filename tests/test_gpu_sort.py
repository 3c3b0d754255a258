"""GPU parity: clo_sort (onesweep radix, bitonic network, gselect, pairs, partition)
vs the CPU oracle, through the C-ABI.  Bit-exact, stable order included."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

RADIX_SIZES = [1, 2, 3, 100, 4096, 8191, 8192, 8193, 65536, 100003, (1 << 20) + 5]


def _rand(rng, ctype, n):
    dt = oracle.NP_TYPES[ctype]
    if np.issubdtype(dt, np.integer):
        info = np.iinfo(dt)
        return rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
    return ((rng.random(n) - 0.5) * 2000).astype(dt)


@pytest.mark.parametrize("n", RADIX_SIZES)
@pytest.mark.parametrize("et", [oracle.UINT, oracle.ULONG, oracle.UCHAR, oracle.USHORT, oracle.INT, oracle.LONG])
def test_satradix_keys(clo, ctx, queue, et, n):
    rng = np.random.default_rng(n + 13 * et)
    a = _rand(rng, et, n)
    s = clo.CloSort("satradix", ctx, et)
    got = s.with_host_data(a, queue)
    s.destroy()
    want = oracle.sort_satradix(a, et, radix=16, lws=64)
    assert np.array_equal(got, want)
    # raw-bit ascending order == numpy sort of the unsigned view
    u = a.view(np.dtype("u%d" % a.dtype.itemsize))
    assert np.array_equal(got.view(u.dtype), np.sort(u))


@pytest.mark.parametrize("opts", ["radix=2", "radix=4", "radix=256", "radix=16,scan=blelloch", "radix=32"])
def test_satradix_radix_option_results(clo, ctx, queue, opts):
    a = oracle.sort_input(0, oracle.UINT, 1 << 14)
    s = clo.CloSort("satradix", ctx, oracle.UINT, options=opts)
    got = s.with_host_data(a, queue)
    s.destroy()
    radix = int(opts.split(",")[0].split("=")[1])
    want = oracle.sort_satradix(a, oracle.UINT, radix=radix, lws=256)
    assert np.array_equal(got, want)


def test_satradix_skewed_and_constant(clo, ctx, queue):
    rng = np.random.default_rng(5)
    n = (1 << 18) + 77
    for a in (np.zeros(n, np.uint32), np.full(n, 0xFFFFFFFF, np.uint32),
              (rng.zipf(1.3, n) % 1000).astype(np.uint32), np.arange(n, dtype=np.uint32)[::-1].copy()):
        s = clo.CloSort("satradix", ctx, oracle.UINT)
        got = s.with_host_data(a, queue)
        s.destroy()
        assert np.array_equal(got, np.sort(a))


@pytest.mark.parametrize("get_key,shift,mask,kt", [("((x) >> 32)", 32, None, oracle.UINT),
                                                   ("((x) & 0xFFFF)", 0, 0xFFFF, oracle.UINT),
                                                   ("(((x) >> 8) & 0xFF)", 8, 0xFF, oracle.UCHAR)])
def test_satradix_packed_key_value_is_stable(clo, ctx, queue, get_key, shift, mask, kt):
    """elem = ulong packed (key | payload); equal keys must keep their input order."""
    rng = np.random.default_rng(11)
    n = 70001
    a = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    a = (a & ~np.uint64(0xFFFF00000000)) | (rng.integers(0, 50, size=n).astype(np.uint64) << np.uint64(32))
    s = clo.CloSort("satradix", ctx, oracle.ULONG, key_type=kt, get_key=get_key)
    got = s.with_host_data(a, queue)
    s.destroy()
    want = oracle.sort_satradix(a, oracle.ULONG, radix=16, lws=64, key_type=kt, shift=shift, mask=mask)
    assert np.array_equal(got, want)


def test_satradix_device_data_out_of_place_and_reuse(clo, ctx, queue):
    import torch
    n = (1 << 20) + 999
    a = oracle.sort_input(1, oracle.UINT, n)
    t_in = torch.from_numpy(a.view(np.int32)).cuda()
    t_out = torch.empty_like(t_in)
    torch.cuda.synchronize()
    b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
    s = clo.CloSort("satradix", ctx, oracle.UINT)
    s.with_device_data(queue, b_in, b_out, n)
    queue.finish()
    assert np.array_equal(t_in.cpu().numpy().view(np.uint32), a), "input buffer was modified"
    assert np.array_equal(t_out.cpu().numpy().view(np.uint32), np.sort(a))
    for _ in range(2):  # in place, same sorter object twice
        t_in.copy_(torch.from_numpy(a.view(np.int32)))
        torch.cuda.synchronize()
        s.with_device_data(queue, b_in, None, n)
        queue.finish()
        assert np.array_equal(t_in.cpu().numpy().view(np.uint32), np.sort(a))
    b_in.destroy(); b_out.destroy(); s.destroy()


@pytest.mark.parametrize("kt,n", [(oracle.UINT, 100003), (oracle.ULONG, 100003), (oracle.UINT, (1 << 20) + 1),
                                  (oracle.ULONG, 1 << 20)])
def test_sort_pairs_stable(clo, ctx, queue, kt, n):
    import torch
    rng = np.random.default_rng(n + kt)
    dt = oracle.NP_TYPES[kt]
    keys = rng.integers(0, 1000, size=n).astype(dt)            # many duplicates
    keys[::7] = rng.integers(0, np.iinfo(dt).max, size=keys[::7].size, dtype=dt)
    payload = np.arange(n, dtype=np.uint32)
    tk = torch.from_numpy(keys.view(np.int32 if kt == oracle.UINT else np.int64)).cuda()
    tp = torch.from_numpy(payload.view(np.int32)).cuda()
    torch.cuda.synchronize()
    bk, bp = clo.Buffer.wrap_tensor(ctx, tk), clo.Buffer.wrap_tensor(ctx, tp)
    s = clo.CloSort("satradix", ctx, kt)
    s.pairs_with_device_data(queue, bk, bp, n)
    queue.finish()
    wk, wp = oracle.sort_pairs(keys, payload, kt)
    assert np.array_equal(tk.cpu().numpy().view(dt), wk)
    assert np.array_equal(tp.cpu().numpy().view(np.uint32), wp)
    bk.destroy(); bp.destroy(); s.destroy()


@pytest.mark.parametrize("alg", ["sbitonic", "abitonic"])
@pytest.mark.parametrize("n", [2, 4, 64, 1024, 4096, 8192, 1 << 16, 1 << 20])
def test_bitonic_power_of_two(clo, ctx, queue, alg, n):
    a = oracle.sort_input(0, oracle.UINT, n)
    s = clo.CloSort(alg, ctx, oracle.UINT)
    got = s.with_host_data(a, queue)
    s.destroy()
    assert np.array_equal(got, oracle.sort_bitonic(a, oracle.UINT))


@pytest.mark.parametrize("et", [oracle.CHAR, oracle.UCHAR, oracle.SHORT, oracle.USHORT, oracle.INT, oracle.LONG,
                                oracle.ULONG, oracle.FLOAT, oracle.DOUBLE])
@pytest.mark.parametrize("desc", [False, True])
def test_bitonic_types_and_compare(clo, ctx, queue, et, desc):
    rng = np.random.default_rng(et)
    a = _rand(rng, et, 1 << 13)
    s = clo.CloSort("sbitonic", ctx, et, compare="((a) < (b))" if desc else None)
    got = s.with_host_data(a, queue)
    s.destroy()
    want = oracle.sort_bitonic(a, et, descending=desc)
    assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    ref = np.sort(a)[::-1] if desc else np.sort(a)
    assert np.array_equal(got, ref)


def test_bitonic_network_with_get_key_is_bit_exact(clo, ctx, queue):
    """unstable network + non-trivial key: only the canonical network reproduces this."""
    rng = np.random.default_rng(3)
    n = 1 << 14
    a = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
    a = (a & ~np.uint64(0xFF)) | rng.integers(0, 4, size=n).astype(np.uint64)
    s = clo.CloSort("abitonic", ctx, oracle.ULONG, key_type=oracle.UCHAR, get_key="((x) & 0xFF)")
    got = s.with_host_data(a, queue)
    s.destroy()
    want = oracle.sort_bitonic(a, oracle.ULONG, key_type=oracle.UCHAR, mask=0xFF)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [3, 100, 5000, 4097, 100003])
def test_bitonic_any_n(clo, ctx, queue, n):
    a = oracle.sort_input(2, oracle.UINT, n)
    s = clo.CloSort("sbitonic", ctx, oracle.UINT)
    got = s.with_host_data(a, queue)
    s.destroy()
    assert np.array_equal(got, np.sort(a))


@pytest.mark.parametrize("n", [1, 17, 256, 1000, 5000])
def test_gselect_stable_rank(clo, ctx, queue, n):
    rng = np.random.default_rng(n)
    a = (rng.integers(0, 1 << 62, size=n, dtype=np.uint64) & ~np.uint64(0xFF)) | \
        rng.integers(0, 8, size=n).astype(np.uint64)
    s = clo.CloSort("gselect", ctx, oracle.ULONG, key_type=oracle.UCHAR, get_key="((x) & 0xFF)")
    got = s.with_host_data(a, queue)
    s.destroy()
    assert np.array_equal(got, oracle.sort_gselect(a, oracle.ULONG, key_type=oracle.UCHAR, mask=0xFF))


def test_partition_by_splitters(clo, ctx, queue):
    """sample-sort building block: stable multi-way partition with (key, index) splitters."""
    import torch
    rng = np.random.default_rng(9)
    n, P, g0 = 300001, 8, 10_000_000
    keys = rng.integers(0, 50, size=n).astype(np.uint32)      # heavy duplication
    payload = np.arange(n, dtype=np.uint32)
    sk = np.array([5, 5, 20, 20, 20, 33, 49], dtype=np.uint32)
    si = np.array([g0 + 10, g0 + 200000, g0 + 5, g0 + 100000, g0 + 250000, 0, g0 + n], dtype=np.uint64)
    g = g0 + np.arange(n, dtype=np.uint64)
    bucket = np.zeros(n, dtype=np.int64)
    for k_, i_ in zip(sk, si):
        bucket += ((k_ < keys) | ((k_ == keys) & (i_ <= g))).astype(np.int64)
    order = np.argsort(bucket, kind="stable")
    t = {name: torch.from_numpy(arr).cuda() for name, arr in
         dict(k=keys.view(np.int32), p=payload.view(np.int32), sk=sk.view(np.int32), si=si.view(np.int64)).items()}
    t["ko"], t["po"] = torch.empty_like(t["k"]), torch.empty_like(t["p"])
    t["cnt"] = torch.zeros(P, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    b = {name: clo.Buffer.wrap_tensor(ctx, x) for name, x in t.items()}
    s = clo.CloSort("satradix", ctx, oracle.UINT)
    s.partition_with_device_data(queue, b["k"], b["p"], b["ko"], b["po"], n, g0, b["sk"], b["si"], P, b["cnt"])
    queue.finish()
    assert np.array_equal(t["cnt"].cpu().numpy(), np.bincount(bucket, minlength=P))
    assert np.array_equal(t["ko"].cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(t["po"].cpu().numpy().view(np.uint32), payload[order])
    for x in b.values():
        x.destroy()
    s.destroy()


def test_sort_errors(clo, ctx):
    with pytest.raises(clo.CloError) as ei:
        clo.CloSort("quicksort", ctx, oracle.UINT)
    assert ei.value.code == clo.CLO_ERROR_IMPL_NOT_FOUND
    for kw in (dict(options="radix=3"), dict(options="bogus=1"), dict(options="radix")):
        with pytest.raises(clo.CloError) as ei:
            clo.CloSort("satradix", ctx, oracle.UINT, **kw)
        assert ei.value.code == clo.CLO_ERROR_ARGS
    for kw in (dict(options="minps=5"), dict(options="minps=3,maxps=2"), dict(options="zzz=1")):
        with pytest.raises(clo.CloError) as ei:
            clo.CloSort("abitonic", ctx, oracle.UINT, **kw)
        assert ei.value.code == clo.CLO_ERROR_ARGS
    # the radix kernels never expand CLO_SORT_COMPARE (clo_sort_satradix.cl defines no comparison):
    # any compare string is accepted and ignored; a get_key that does not compile is an error
    clo.CloSort("satradix", ctx, oracle.UINT, compare="((a) >= (b))").destroy()
    with pytest.raises(clo.CloError) as ei:
        clo.CloSort("satradix", ctx, oracle.UINT, get_key="((x) >>> 3)")
    assert ei.value.code == clo.CLO_ERROR_ARGS
    s = clo.CloSort("satradix", ctx, oracle.UINT)
    assert s.kernel_names() == ["clo_radix_histogram", "clo_radix_scan_bins", "clo_radix_onesweep_v6"]
    s.destroy()


@pytest.mark.parametrize("et", [oracle.UINT, oracle.ULONG])
def test_satradix_atomic_rank_repair_path(clo, ctx, queue, et, monkeypatch):
    """The keys-only onesweep kernel ranks with shared-memory atomics and VERIFIES every tile it
    writes; CLO_RADIX_PP_FLAGS=8 makes every fifth tile be written wrong and reported, so the
    result is only sorted if the ballot-rank repair path works."""
    monkeypatch.setenv("CLO_RADIX_PP_FLAGS", "8")
    rng = np.random.default_rng(77 + et)
    n = (1 << 20) + 4321
    a = _rand(rng, et, n)
    a[: n // 3] &= 0xFFFF                      # duplicates in the upper digits
    s = clo.CloSort("satradix", ctx, et)
    got = s.with_host_data(a, queue)
    dbg = s.debug(queue)
    s.destroy()
    monkeypatch.delenv("CLO_RADIX_PP_FLAGS")
    clo.CloSort("satradix", ctx, et).destroy()   # re-reads the environment: flag off again
    assert np.array_equal(got, np.sort(a))
    assert dbg[0] == 0 and dbg[1] > 0          # no timeout; tiles were repaired


@pytest.mark.parametrize("et,P,with_payload", [(oracle.UINT, 2, False), (oracle.ULONG, 4, True),
                                               (oracle.UINT, 16, True), (oracle.ULONG, 3, False)])
def test_partition_stages_scatter_to_destinations(clo, ctx, queue, et, P, with_payload):
    """Fused partition + exchange building blocks: stage 1 counts, stage 2 scatters bucket q to
    dests[q] + first_slot[q] (here: one tensor per bucket, standing in for the peers' receive
    buffers).  Stable inside every bucket; *ok == 0 writes nothing."""
    import torch
    rng = np.random.default_rng(100 + P)
    dt = oracle.NP_TYPES[et]
    n, g0 = 200003, 5_000_000_000
    keys = rng.integers(0, 40, size=n).astype(dt)
    payload = np.arange(n, dtype=np.uint32)
    sk = np.sort(rng.integers(0, 40, size=P - 1)).astype(dt)
    si = (g0 + np.sort(rng.integers(0, n, size=P - 1))).astype(np.uint64)
    o = np.lexsort((si, sk)); sk, si = sk[o], si[o]
    g = g0 + np.arange(n, dtype=np.uint64)
    bucket = np.zeros(n, dtype=np.int64)
    for k_, i_ in zip(sk, si):
        bucket += ((k_ < keys) | ((k_ == keys) & (i_ <= g))).astype(np.int64)
    tdt = torch.int32 if dt().itemsize == 4 else torch.int64
    sdt = np.int32 if dt().itemsize == 4 else np.int64
    tk = torch.from_numpy(keys.view(sdt)).cuda()
    tp = torch.from_numpy(payload.view(np.int32)).cuda()
    tsk = torch.from_numpy(sk.view(sdt)).cuda()
    tsi = torch.from_numpy(si.view(np.int64)).cuda()
    cnt = torch.zeros(P, dtype=torch.int64, device="cuda")
    s = clo.CloSort("satradix", ctx, et)
    W = clo.Buffer.wrap_tensor
    b = [W(ctx, tk), W(ctx, tsk), W(ctx, tsi), W(ctx, cnt), W(ctx, tp)]
    s.partition_count_with_device_data(queue, b[0], n, g0, b[1], b[2], P, b[3])
    queue.finish()
    want_cnt = np.bincount(bucket, minlength=P)
    assert np.array_equal(cnt.cpu().numpy(), want_cnt)
    lead = [int(x) for x in rng.integers(0, 100, size=P)]          # unaligned first slots
    dk = [torch.full((lead[q] + int(want_cnt[q]) + 64,), -1, dtype=tdt, device="cuda") for q in range(P)]
    dp = [torch.full((lead[q] + int(want_cnt[q]) + 64,), -1, dtype=torch.int32, device="cuda") for q in range(P)]
    ptr_k = torch.zeros(16, dtype=torch.int64); ptr_p = torch.zeros(16, dtype=torch.int64)
    for q in range(P):
        ptr_k[q], ptr_p[q] = dk[q].data_ptr(), dp[q].data_ptr()
    ptr_k, ptr_p = ptr_k.cuda(), ptr_p.cuda()
    fs = torch.zeros(16, dtype=torch.int64); fs[:P] = torch.tensor(lead); fs = fs.cuda()
    for okv in (0, 1):
        ok = torch.tensor([okv], dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        bb = [W(ctx, fs), W(ctx, ptr_k), W(ctx, ptr_p), W(ctx, ok)]
        s.partition_scatter_with_device_data(queue, b[0], b[4] if with_payload else None, n, g0, b[1], b[2], P,
                                             bb[0], bb[1], bb[2] if with_payload else None, bb[3])
        queue.finish()
        for x in bb:
            x.destroy()
        for q in range(P):
            got = dk[q].cpu().numpy().view(dt)
            c = int(want_cnt[q])
            if okv == 0:
                assert np.all(dk[q].cpu().numpy() == -1)                  # no-op
                continue
            assert np.array_equal(got[lead[q]:lead[q] + c], keys[bucket == q])    # stable
            assert np.all(dk[q].cpu().numpy()[:lead[q]] == -1) and np.all(dk[q].cpu().numpy()[lead[q] + c:] == -1)
            if with_payload:
                assert np.array_equal(dp[q].cpu().numpy().view(np.uint32)[lead[q]:lead[q] + c], payload[bucket == q])
    for x in b:
        x.destroy()
    s.destroy()


def test_sample_sort_two_gpus_fused_exchange():
    """2 ranks, one per GPU: fused partition + peer-memory exchange + local sort equals the
    stable sort of the concatenated input (skipped on a one-GPU box)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST GPU OK" in r.stdout


def test_c_client_dist_sort():
    """tools/dist_sort_nccl.c: a plain C program (one forked process per GPU, NCCL for the three
    callbacks) sorts and scans over the GPUs of the box through clo_dist_* and checks itself --
    the multi-GPU layer is reachable without Python.  With one GPU the world is 1."""
    import json
    import os
    import subprocess
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tools", "dist_sort_nccl")
    if not os.path.exists(exe):
        pytest.skip("tools/dist_sort_nccl not built (needs nccl.h at build time)")
    gpus = min(2, torch.cuda.device_count())
    r = subprocess.run([exe, str(gpus), "22"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["ok"] is True and line["gpus"] == gpus


@pytest.mark.parametrize("et,with_payload", [(oracle.UINT, False), (oracle.ULONG, False), (oracle.UINT, True), (oracle.ULONG, True)])
def test_satradix_wide_lookback_words(clo, ctx, queue, et, with_payload, monkeypatch):
    """n >= 2^31 switches the tile-prefix words to 64 bits; CLO_RADIX_WIDE=1 forces that path at
    a testable size (keys only and keys + payload, stable)."""
    import torch
    monkeypatch.setenv("CLO_RADIX_WIDE", "1")
    rng = np.random.default_rng(5 + et)
    n = (1 << 20) + 999
    a = _rand(rng, et, n)
    a[: n // 2] &= 0xFF                        # many equal keys: stability matters
    s = clo.CloSort("satradix", ctx, et)
    if not with_payload:
        got = s.with_host_data(a, queue)
        assert np.array_equal(got, np.sort(a))
    else:
        sdt = np.int32 if a.dtype.itemsize == 4 else np.int64
        tk = torch.from_numpy(a.view(sdt).copy()).cuda()
        tp = torch.arange(n, dtype=torch.int32, device="cuda")
        bk, bp = clo.Buffer.wrap_tensor(ctx, tk), clo.Buffer.wrap_tensor(ctx, tp)
        s.pairs_with_device_data(queue, bk, bp, n)
        queue.finish()
        order = np.argsort(a, kind="stable")
        assert np.array_equal(tk.cpu().numpy().view(a.dtype), a[order])
        assert np.array_equal(tp.cpu().numpy().view(np.uint32), order.astype(np.uint32))
        bk.destroy(); bp.destroy()
    dbg = s.debug(queue)
    s.destroy()
    monkeypatch.delenv("CLO_RADIX_WIDE")
    clo.CloSort("satradix", ctx, et).destroy()
    assert dbg[0] == 0


def _np_bitonic_network(a, key, after):
    """The canonical network (clo_sort_sbitonic.cl:38-69 + host loop clo_sort_sbitonic.c:73-118)
    in numpy, for an arbitrary key(x) and after(ka, kb); n padded to a power of two with
    elements that come after everything."""
    n = len(a)
    np2 = 1
    while np2 < n:
        np2 <<= 1
    d = np.concatenate([a, np.zeros(np2 - n, dtype=a.dtype)])
    pad = np.concatenate([np.zeros(n, dtype=bool), np.ones(np2 - n, dtype=bool)])
    log = int(np.log2(np2)) if np2 > 1 else 0
    p = np.arange(np2 // 2)
    for stage in range(1, log + 1):
        desc = ((p >> (stage - 1)) & 1).astype(bool)
        for step in range(stage, 0, -1):
            stride = 1 << (step - 1)
            i1 = p + (p // stride) * stride
            i2 = i1 + stride
            c = after(key(d[i1]), key(d[i2]))
            must = (pad[i1] & ~pad[i2]) | ((pad[i1] == pad[i2]) & c)
            sw = must != desc
            x, y = d[i1][sw].copy(), d[i2][sw].copy()
            d[i1[sw]], d[i2[sw]] = y, x
            px, py = pad[i1][sw].copy(), pad[i2][sw].copy()
            pad[i1[sw]], pad[i2[sw]] = py, px
    return d[:n]


@pytest.mark.parametrize("alg", ["sbitonic", "abitonic"])
@pytest.mark.parametrize("n", [1, 2, 1000, 4096, 5000])
def test_custom_macro_strings_are_compiled_at_run_time(clo, ctx, queue, alg, n):
    """compare / get_key strings outside the precompiled menu: spliced into the network and built
    with NVRTC, as the reference builds them into its OpenCL kernels (clo_sort_abstract.c:144-168)."""
    rng = np.random.default_rng(n)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    s = clo.CloSort(alg, ctx, oracle.UINT, key_type=oracle.UCHAR,
                    compare="(((a) & 15) > ((b) & 15))", get_key="(((x) >> 3) ^ (x))")
    got = s.with_host_data(a, queue)
    s.destroy()
    key = lambda x: ((x >> np.uint32(3)) ^ x).astype(np.uint8)
    after = lambda ka, kb: (ka & 15) > (kb & 15)
    assert np.array_equal(got, _np_bitonic_network(a, key, after))


def test_custom_macro_strings_gselect_and_errors(clo, ctx, queue):
    rng = np.random.default_rng(3)
    a = rng.integers(0, 1000, size=3001).astype(np.uint32)
    # gselect ranks by COMPARE plus KEY equality (clo_sort_gselect.cl:44-56), so the comparator must be a
    # strict order on the keys: descending by the scrambled key, equal elements in input order
    s = clo.CloSort("gselect", ctx, oracle.UINT, compare="((a) < (b))", get_key="((x) * 40503u % 65521u)")
    got = s.with_host_data(a, queue)
    s.destroy()
    k = (a.astype(np.uint64) * 40503 % (1 << 32) % 65521).astype(np.int64)
    order = np.argsort(-k, kind="stable")
    assert np.array_equal(got, a[order])
    # a float element with a custom comparator (descending by magnitude)
    f = ((rng.random(2048) - 0.5) * 100).astype(np.float32)
    s = clo.CloSort("sbitonic", ctx, oracle.FLOAT, compare="(fabsf(a) < fabsf(b))")
    got = s.with_host_data(f, queue)
    s.destroy()
    assert np.array_equal(got, _np_bitonic_network(f, lambda x: x, lambda ka, kb: np.abs(ka) < np.abs(kb)))
    # strings that do not compile are an argument error carrying the compiler's message
    with pytest.raises(clo.CloError) as ei:
        clo.CloSort("sbitonic", ctx, oracle.UINT, compare="((a) >>> (b))")
    assert ei.value.code == clo.CLO_ERROR_ARGS and "do not compile" in str(ei.value)


@pytest.mark.parametrize("et,kt,get_key,keyfn", [
    (oracle.UINT, oracle.UINT, "((x) * 40503u)", lambda x: (x.astype(np.uint64) * 40503 % (1 << 32)).astype(np.uint64)),
    (oracle.ULONG, oracle.USHORT, "(((x) >> 7) ^ (x))", lambda x: ((x >> np.uint64(7)) ^ x) & np.uint64(0xFFFF)),
    (oracle.INT, oracle.CHAR, "((x) / 3)", lambda x: (np.trunc(x.astype(np.int64) / 3).astype(np.int64).astype(np.int8)).view(np.uint8).astype(np.uint64)),
    (oracle.USHORT, oracle.UINT, "(~(x))", lambda x: (~x.astype(np.uint32)).astype(np.uint64) & np.uint64(0xFFFF)),
])
@pytest.mark.parametrize("n", [1, 777, (1 << 18) + 11])
def test_satradix_custom_get_key_is_compiled_at_run_time(clo, ctx, queue, et, kt, get_key, keyfn, n):
    """satradix with a get_key string outside the menu (clo_sort_abstract.c:157-168 splices any macro
    body; clo_sort_satradix.cl:58-61 cuts the digit from `CLO_SORT_KEY_GET(value)`): the key
    extraction is built with NVRTC and the result is the stable sort by the key's raw bits -- the
    low elem-size bits of them, as the reference's host loop only walks those
    (clo_sort_satradix.c:166-169).  In place and out of place."""
    import torch
    rng = np.random.default_rng(n + et)
    a = _rand(rng, et, n)
    k = keyfn(a)
    eb = 8 * a.dtype.itemsize
    if eb < 64:
        k = k & np.uint64((1 << eb) - 1)
    want = a[np.argsort(k, kind="stable")]
    s = clo.CloSort("satradix", ctx, et, key_type=kt, get_key=get_key)
    got = s.with_host_data(a, queue)
    assert np.array_equal(got, want)
    sdt = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[a.dtype.itemsize]
    t = torch.from_numpy(a.view(sdt).copy()).cuda()
    b = clo.Buffer.wrap_tensor(ctx, t)
    s.with_device_data(queue, b, None, n)          # in place
    queue.finish()
    assert np.array_equal(t.cpu().numpy().view(a.dtype), want)
    b.destroy()
    s.destroy()


@pytest.mark.parametrize("alg", ["sbitonic", "gselect"])
def test_integer_elements_with_a_float_key_are_converted_by_value(clo, ctx, queue, alg):
    """elem uint, key float, default get_key: the reference's kernels compute `(float) (x)` -- a
    VALUE conversion (clo_sort_abstract.c:157-168) -- so values >= 2^31 must sort as large
    numbers, not as negative IEEE patterns.  These type mixes go through the run-time compiler."""
    rng = np.random.default_rng(9)
    a = rng.integers(0, 2**32, size=4096, dtype=np.uint64).astype(np.uint32)
    a[:8] = [0xFFFFFFFF, 0x80000000, 0x7FFFFFFF, 0, 1, 0xC0000000, 0x3F800000, 0xBF800000]
    s = clo.CloSort(alg, ctx, oracle.UINT, key_type=oracle.FLOAT)
    got = s.with_host_data(a, queue)
    s.destroy()
    assert np.array_equal(np.sort(got), np.sort(a))                 # a permutation
    k = got.astype(np.float32)                                      # the key the kernels compared
    assert np.all(k[1:] >= k[:-1])
    # and the reverse: float elements ordered by their truncated integer value
    f = ((rng.random(2048) - 0.5) * 1000).astype(np.float32)
    s = clo.CloSort(alg, ctx, oracle.FLOAT, key_type=oracle.INT)
    got = s.with_host_data(f, queue)
    s.destroy()
    assert np.array_equal(np.sort(got), np.sort(f))
    ki = np.trunc(got).astype(np.int64)
    assert np.all(ki[1:] >= ki[:-1])


@pytest.mark.parametrize("et", [oracle.CHAR, oracle.SHORT, oracle.INT, oracle.LONG, oracle.FLOAT, oracle.DOUBLE])
@pytest.mark.parametrize("n", [1, 1000, (1 << 18) + 77])
def test_satradix_typed_order_opt_in(clo, ctx, queue, et, n):
    """options "typed_order=1" (an opt-in deviation): signed integers and floats in
    numeric order.  Without it the reference's raw-bit order stays (negative ints after positive
    ones) and float keys are rejected, as `key >> b` does not compile for them."""
    rng = np.random.default_rng(n + et)
    a = _rand(rng, et, n)
    if a.dtype.kind == "f" and n >= 8:
        a[:4] = [np.inf, -np.inf, 0.0, np.finfo(a.dtype).tiny]
    s = clo.CloSort("satradix", ctx, et, options="typed_order=1")
    got = s.with_host_data(a, queue)
    s.destroy()
    assert np.array_equal(got, np.sort(a))
    if a.dtype.kind == "i":
        s = clo.CloSort("satradix", ctx, et)                       # the reference's order: raw bits
        got = s.with_host_data(a, queue)
        s.destroy()
        u = a.view({1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[a.dtype.itemsize])
        assert np.array_equal(got.view(u.dtype), np.sort(u))
    else:
        s = clo.CloSort("satradix", ctx, et)
        with pytest.raises(clo.CloError):
            s.with_host_data(a, queue)
        s.destroy()


@pytest.mark.parametrize("et", [oracle.UINT, oracle.ULONG])
@pytest.mark.parametrize("shape", ["low16", "low24", "top8", "equal", "mid", "bit0"])
@pytest.mark.parametrize("with_payload", [False, True])
def test_satradix_identity_passes_move_nothing(clo, ctx, queue, et, shape, with_payload):
    """Passes whose digit is the same for every key are skipped on the device (the passes hand the
    location of the keys on through chain words, one conditional copy at the end): any subset of
    the passes can be an identity -- leading, trailing, in the middle, all of them -- in place and
    out of place, keys only and with a payload (stable)."""
    import torch
    rng = np.random.default_rng(hash(shape) % 1000 + et)
    n = (1 << 18) + 321
    bits = 8 * oracle.NP_TYPES[et]().itemsize
    a = _rand(rng, et, n)
    mask = {"low16": 0xFFFF, "low24": 0xFFFFFF, "top8": 0xFF << (bits - 8), "equal": 0, "mid": 0xFF00 << 8, "bit0": 1}[shape]
    a = (a & a.dtype.type(mask)) | a.dtype.type(0x0100000000000000 if (bits == 64 and shape != "top8") else 0)
    s = clo.CloSort("satradix", ctx, et)
    sdt = np.int32 if a.dtype.itemsize == 4 else np.int64
    if not with_payload:
        got = s.with_host_data(a, queue)                      # out of place on the device
        assert np.array_equal(got, np.sort(a))
        t = torch.from_numpy(a.view(sdt).copy()).cuda()
        torch.cuda.synchronize()
        b = clo.Buffer.wrap_tensor(ctx, t)
        for rep in range(2):                                  # in place, twice (second time already sorted)
            s.with_device_data(queue, b, None, n)
            queue.finish()
            assert np.array_equal(t.cpu().numpy().view(a.dtype), np.sort(a))
        b.destroy()
    else:
        tk = torch.from_numpy(a.view(sdt).copy()).cuda()
        tp = torch.arange(n, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        bk, bp = clo.Buffer.wrap_tensor(ctx, tk), clo.Buffer.wrap_tensor(ctx, tp)
        s.pairs_with_device_data(queue, bk, bp, n)
        queue.finish()
        order = np.argsort(a, kind="stable")
        assert np.array_equal(tk.cpu().numpy().view(a.dtype), a[order])
        assert np.array_equal(tp.cpu().numpy(), order.astype(np.int32))
        bk.destroy(); bp.destroy()
    s.destroy()


def test_introspection_getters_name_real_kernels(clo, ctx):
    """clo_sort_get_kernel_name / clo_scan_get_kernel_name (clo_sort_abstract.c:571-629) return the
    names of CUDA kernels that exist in the library, the counts of the per-algorithm headers, and
    non-trivial shared-memory sizes."""
    import ctypes, os, re
    L = clo.lib()
    blob = open(os.path.join(os.path.dirname(clo.__file__), "libcl_ops.so"), "rb").read()
    hdr = os.path.join(os.path.dirname(os.path.dirname(clo.__file__)), "include", "cl_ops")
    want = {"sbitonic": "CLO_SORT_SBITONIC_NUM_KERNELS", "abitonic": "CLO_SORT_ABITONIC_NUM_KERNELS",
            "gselect": "CLO_SORT_GSELECT_NUM_KERNELS", "satradix": "CLO_SORT_SATRADIX_NUM_KERNELS"}
    L.clo_sort_get_localmem_usage.restype = ctypes.c_size_t
    for alg, macro in want.items():
        text = open(os.path.join(hdr, "clo_sort_%s.h" % alg)).read()
        n_hdr = int(re.search(r"#define\s+%s\s+(\d+)" % macro, text).group(1))
        s = clo.CloSort(alg, ctx, clo.UINT)
        names = s.kernel_names()
        assert len(names) == n_hdr
        for i, nm in enumerate(names):
            assert nm.encode() in blob, "%s is not a kernel of the library" % nm
            assert ('"%s"' % nm) in text, "%s is not named by clo_sort_%s.h" % (nm, alg)
            e = clo._Err()
            L.clo_sort_get_localmem_usage(s.h, i, 0, 1 << 20, e.ref())
            e.check()
        s.destroy()
    sc = clo.CloScan("blelloch", ctx, clo.UINT, clo.UINT)
    L.clo_scan_get_kernel_name.restype = ctypes.c_char_p
    e = clo._Err()
    n = L.clo_scan_get_num_kernels(sc.h, e.ref())
    text = open(os.path.join(hdr, "clo_scan_blelloch.h")).read()
    assert n == int(re.search(r"#define\s+CLO_SCAN_BLELLOCH_NUM_KERNELS\s+(\d+)", text).group(1))
    for i in range(n):
        nm = L.clo_scan_get_kernel_name(sc.h, i, e.ref())
        assert nm in blob and ('"%s"' % nm.decode()) in text
    sc.destroy()
