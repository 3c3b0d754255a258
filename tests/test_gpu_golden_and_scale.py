"""GPU parity, two ends of the scale:

* the committed golden fixtures (tests/golden/ref_{sort,scan,rng}.npz, produced by the
  reference's own kernels, see make_golden.py) compared DIRECTLY with the CUDA path through the
  C-ABI -- not only through the oracle;
* the BASELINE.json sizes: a 2^28-key satradix sort against the OpenMP oracle, and 2^32 RNG words
  through a checksum of checksums (per-stream xor/sum against the oracle on a slice, layout
  property on the whole buffer);
* the hooks that make the safety nets of the onesweep pass testable: a planted inversion in front
  of the stability detector, and a reported prefix time-out.
"""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CT = {"uint": oracle.UINT, "int": oracle.INT, "ulong": oracle.ULONG, "long": oracle.LONG,
      "uchar": oracle.UCHAR, "ushort": oracle.USHORT, "float": oracle.FLOAT}
HASH = {"none": None, "knuth": "KNUTH(x)", "xs1": "XS1(x)"}
# golden variant -> clo_sort_new arguments (element type, key type, get_key, compare)
SORT_VARIANTS = {
    "uint": ("uint", None, None, None), "uint_desc": ("uint", None, None, "((a) < (b))"), "int": ("int", None, None, None),
    "ulong": ("ulong", None, None, None), "uchar": ("uchar", None, None, None), "ushort": ("ushort", None, None, None),
    "float": ("float", None, None, None),
    "ulong_keylo8": ("ulong", "uchar", "((x) & 0xFF)", None),
    "ulong_keyhi32": ("ulong", "uint", "((x) >> 32)", None),
}


def _cases(gold):
    return sorted({k.rsplit("/", 1)[0] for k in gold.files})


def test_golden_sort_fixtures_directly(clo, ctx, queue):
    gold = np.load(os.path.join(G, "ref_sort.npz"))
    seen = {"sbitonic": 0, "gselect": 0, "satradix": 0}
    for case in _cases(gold):
        parts = case.split("/")
        alg, name = parts[0], parts[1]
        elem, key, get_key, compare = SORT_VARIANTS[name]
        a, want = gold[case + "/in"], gold[case + "/out"]
        opts = None
        if alg.startswith("satradix"):
            if name in ("float", "uint_desc"):
                continue                      # satradix ignores compare and rejects float keys (as the reference)
            opts = "radix=%d" % int(alg[len("satradix"):])
            alg = "satradix"
        s = clo.CloSort(alg, ctx, CT[elem], key_type=CT[key] if key else None, options=opts, compare=compare, get_key=get_key)
        got = s.with_host_data(a.copy(), queue)
        s.destroy()
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), case
        seen[alg] += 1
    assert seen["sbitonic"] == 27 and seen["gselect"] == 9 and seen["satradix"] >= 30


def test_golden_scan_fixtures_directly(clo, ctx, queue):
    gold = np.load(os.path.join(G, "ref_scan.npz"))
    n_cases = 0
    for case in _cases(gold):
        name = case.split("/")[0]
        e, s_ = name.split("_")
        a, want = gold[case + "/in"], gold[case + "/out"]
        sc = clo.CloScan("blelloch", ctx, CT[e], CT[s_])
        got = sc.with_host_data(a.copy(), queue)
        sc.destroy()
        if e == "float":
            # the reference's own tree order is one of many valid float orders: both sit inside
            # the stated tolerance of a double-precision prefix sum
            ref = oracle.scan_f64ref(a, oracle.FLOAT)
            assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-3), case
            assert np.all(np.abs(want - ref) <= 1e-5 * np.abs(ref) + 1e-3), case
        else:
            assert np.array_equal(got, want), case
        n_cases += 1
    assert n_cases == 30


@pytest.mark.parametrize("rng", oracle.RNG_NAMES)
@pytest.mark.parametrize("h", ["none", "knuth", "xs1"])
def test_golden_rng_fixtures_directly(clo, ctx, queue, rng, h):
    gold = np.load(os.path.join(G, "ref_rng.npz"))
    Gn, runs, main_seed = [int(x) for x in gold["meta"]]
    r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, Gn, main_seed, HASH[h], queue)
    assert np.array_equal(r.read_seeds(queue).view(np.uint8).reshape(-1), gold["%s/%s/seeds" % (rng, h)].view(np.uint8).reshape(-1))
    assert np.array_equal(r.generate_host(runs, queue=queue), gold["%s/%s/out32" % (rng, h)])
    assert np.array_equal(r.read_seeds(queue).view(np.uint8).reshape(-1), gold["%s/%s/states_after" % (rng, h)].view(np.uint8).reshape(-1))
    assert np.array_equal(r.generate_host(3, bits=8, queue=queue), gold["%s/%s/out8_cont" % (rng, h)])
    assert np.array_equal(r.generate_host(3, maxint=1000, queue=queue), gold["%s/%s/outmax1000_cont" % (rng, h)])
    r.destroy()


def test_satradix_2p28_keys_against_the_oracle(clo, ctx, queue):
    """BASELINE.json configs[1] at full size: bit-exact against the oracle's OpenMP satradix."""
    n = 1 << 28
    rng = np.random.default_rng(28)
    a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    s = clo.CloSort("satradix", ctx, oracle.UINT)
    got = s.with_host_data(a, queue)
    dbg = s.debug(queue)
    s.destroy()
    want = oracle.sort_satradix(a, oracle.UINT, radix=16, lws=256, threads=os.cpu_count() or 1)
    assert dbg[0] == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("rng", ["xorshift128", "mwc64x"])
def test_rng_2p32_words_checksum_of_checksums(clo, ctx, queue, rng):
    """BASELINE.json configs[4] at full size (2^22 streams x 2^10 runs = 2^32 words, 16 GiB on the
    device).  Per-run xor and sum over ALL streams, computed on the device, against the oracle's
    per-run xor and sum of the same streams: one word that is wrong anywhere changes them."""
    torch = pytest.importorskip("torch")
    Gn, runs = 1 << 22, 1 << 10
    out = torch.empty(Gn * runs, dtype=torch.int32, device="cuda")
    q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
    bo = clo.Buffer.wrap_tensor(ctx, out)
    r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, Gn, 0, "KNUTH(x)", q)
    r.generate(q, bo, runs)
    torch.cuda.synchronize()
    r.destroy()
    o = out.view(runs, Gn).to(torch.int64) & 0xFFFFFFFF
    dev_sum = o.sum(dim=1).cpu().numpy().astype(np.uint64)
    x = o
    while x.shape[1] > 1:                                  # xor over the stream axis
        h = x.shape[1] // 2
        x = torch.bitwise_xor(x[:, :h], x[:, h:])
    dev_xor = x[:, 0].cpu().numpy().astype(np.uint64)
    del o, x
    # the oracle on the host cores, 2^16 streams per task (ctypes releases the GIL)
    from concurrent.futures import ThreadPoolExecutor
    CH = 1 << 16

    def part(g0):
        seeds = oracle.rng_seeds_dev_gid(rng, 1, 0, CH, gid0=g0)
        w, _ = oracle.rng_generate(rng, seeds, CH, runs)
        w = np.asarray(w).reshape(runs, CH)
        return w.sum(axis=1, dtype=np.uint64), np.bitwise_xor.reduce(w, axis=1).astype(np.uint64)
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        parts = list(ex.map(part, range(0, Gn, CH)))
    want_sum = np.sum([p_[0] for p_ in parts], axis=0, dtype=np.uint64)
    want_xor = np.bitwise_xor.reduce(np.stack([p_[1] for p_ in parts]), axis=0)
    assert np.array_equal(dev_sum, want_sum)
    assert np.array_equal(dev_xor, want_xor)
    bo.destroy(); q.destroy()


@pytest.mark.parametrize("et,pairs", [(oracle.UINT, False), (oracle.ULONG, False), (oracle.ULONG, True)])
def test_stability_detector_sees_a_planted_inversion(clo, ctx, queue, et, pairs, monkeypatch):
    """CLO_RADIX_PP_FLAGS=16 swaps two neighbouring staged entries of one digit BEFORE the
    write-out evaluates its predicate (keys-only: low bits out of order; with a payload: tile
    indices out of order).  The sort is only right if the detector fires and the repair runs."""
    monkeypatch.setenv("CLO_RADIX_PP_FLAGS", "16")
    rng = np.random.default_rng(160 + et)
    n = (1 << 20) + 999
    dt = oracle.NP_TYPES[et]
    a = rng.integers(0, np.iinfo(dt).max, size=n, dtype=dt, endpoint=True)
    a[: n // 2] &= 0xFFFFFF                   # duplicates in the upper digits
    s = clo.CloSort("satradix", ctx, et)
    if pairs:
        torch = pytest.importorskip("torch")
        k = torch.from_numpy(a.view(np.int64)).cuda()
        p = torch.arange(n, dtype=torch.int32, device="cuda")
        q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
        bk, bp = clo.Buffer.wrap_tensor(ctx, k), clo.Buffer.wrap_tensor(ctx, p)
        s.pairs_with_device_data(q, bk, bp, n)
        torch.cuda.synchronize()
        dbg = s.debug(q)
        order = np.argsort(a, kind="stable")
        assert np.array_equal(k.cpu().numpy().view(dt), a[order])
        assert np.array_equal(p.cpu().numpy().astype(np.int64), order)      # stable: payload = source index
        bk.destroy(); bp.destroy(); q.destroy()
    else:
        got = s.with_host_data(a, queue)
        dbg = s.debug(queue)
        assert np.array_equal(got, np.sort(a))
    s.destroy()
    monkeypatch.delenv("CLO_RADIX_PP_FLAGS")
    clo.CloSort("satradix", ctx, et).destroy()   # re-reads the environment: flag off again
    assert dbg[0] == 0 and dbg[1] > 0            # no timeout; the planted inversions were found and repaired


def test_satradix_timeout_is_reported_on_the_device_data_path(clo, ctx, monkeypatch):
    """A prefix time-out (a grid that was not resident) used to be visible only to
    clo_sort_with_host_data.  CLO_RADIX_PP_FLAGS=64 raises the device flag; the NEXT call on the
    sorter must fail with a CLO_ERROR_LIBRARY GError, and the one after that must work again."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("CLO_RADIX_PP_FLAGS", "64")
    n = 1 << 18
    q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
    a = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
    o = torch.empty_like(a)
    ba, bo = clo.Buffer.wrap_tensor(ctx, a), clo.Buffer.wrap_tensor(ctx, o)
    s = clo.CloSort("satradix", ctx, oracle.UINT)
    s.with_device_data(q, ba, bo, n)             # returns: the flag is raised on the device
    torch.cuda.synchronize()
    monkeypatch.delenv("CLO_RADIX_PP_FLAGS")
    clo.CloSort("satradix", ctx, oracle.UINT).destroy()
    with pytest.raises(clo.CloError) as ei:
        s.with_device_data(q, ba, bo, n)
    assert "timed out" in str(ei.value)
    s.with_device_data(q, ba, bo, n)             # reported once; the sorter is usable again
    torch.cuda.synchronize()
    u = o.to(torch.int64) & 0xFFFFFFFF
    assert bool((u[1:] >= u[:-1]).all().item())
    s.destroy(); ba.destroy(); bo.destroy(); q.destroy()
