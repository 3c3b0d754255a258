"""tools/config_runs.py -- BASELINE.json configs[2..4] at their FULL sizes on one B200, with the
size-independent parity properties the tests use at small sizes:
  C3  key-value sort of 2^30 (u64 key, u32 payload = index): sorted, stable, payload a permutation
  C4  exclusive scan of 2^30 u32 (bit-exact vs torch.cumsum in int64) and f32 (vs f64 sum, tolerance)
  C5  2^32 words of xorshift128 / mwc64x: a 2^20-stream x 16-run slice bit-exact vs the oracle and the
      layout property out[r*G+g]
One JSON line per config (device time, CUDA events)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo
import oracle

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6542.1
ctx = clo.Context(); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
what = (sys.argv[1] if len(sys.argv) > 1 else "c3,c4,c5").split(",")


def timed(fn, it=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


if "c3" in what:
    n = 1 << 30
    for dist_name in ("uniform", "zipf(1.0) over 2^24 values"):
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        if dist_name == "uniform":
            keys0 = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
        else:
            u = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
            rank = torch.exp(u * float(np.log(1 << 24))).to(torch.int64).clamp_(1, 1 << 24); del u
            keys0 = rank * -7046029254386353131; del rank          # 0x9E3779B97F4A7C15 as int64 (wraps mod 2^64)
        keys = keys0.clone(); pay = torch.arange(n, dtype=torch.int32, device="cuda")
        s = clo.CloSort("satradix", ctx, clo.ULONG)
        bk, bp = clo.Buffer.wrap_tensor(ctx, keys), clo.Buffer.wrap_tensor(ctx, pay)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.pairs_with_device_data(q, bk, bp, n); torch.cuda.synchronize()     # warm (sorted input afterwards)
        keys.copy_(keys0); pay.copy_(torch.arange(n, dtype=torch.int32, device="cuda")); torch.cuda.synchronize()
        a.record(); s.pairs_with_device_data(q, bk, bp, n); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        uk = keys ^ (-2**63)
        srt = bool((uk[1:] >= uk[:-1]).all().item())
        eq = uk[1:] == uk[:-1]
        stable = bool((pay[1:][eq] > pay[:-1][eq]).all().item())
        gathered = bool(torch.equal(keys0[pay.to(torch.int64)], keys))        # payload = original index of the key
        dbg = s.debug(q)
        print(json.dumps({"config": "C3 key-value sort 2^30 (u64 key + u32 payload), 1 GPU", "keys": dist_name, "ms": round(ms, 2),
                          "gpairs_per_s": round(n / ms / 1e6, 2), "gbs": round(200.0 * n / ms / 1e6, 1), "frac_of_hbm_peak": round(200.0 * n / ms / 1e6 / PEAK, 3),
                          "sorted": srt, "stable": stable, "payload_is_source_index": gathered, "lookback_timeout": dbg[0]}), flush=True)
        bk.destroy(); bp.destroy(); s.destroy(); del keys, pay, uk, eq, keys0
        torch.cuda.empty_cache()

if "c4" in what:
    n = 1 << 30
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    x = torch.randint(0, 128, (n,), dtype=torch.int32, device="cuda", generator=g)
    for st, sdt, bpe in ((clo.UINT, torch.int32, 8), (clo.ULONG, torch.int64, 12)):
        out = torch.empty(n, dtype=sdt, device="cuda")
        sc = clo.CloScan("blelloch", ctx, clo.UINT, st)
        bi, bo = clo.Buffer.wrap_tensor(ctx, x), clo.Buffer.wrap_tensor(ctx, out)
        ms = timed(lambda: sc.with_device_data(q, bi, bo, n))
        ref = torch.cumsum(x.to(torch.int64), 0) - x
        ok = bool(torch.equal(out.to(torch.int64) & (0xFFFFFFFF if sdt == torch.int32 else -1), ref & (0xFFFFFFFF if sdt == torch.int32 else -1)))
        print(json.dumps({"config": "C4 exclusive scan 2^30 u32 -> %s, 1 GPU" % ("u32" if sdt == torch.int32 else "u64"), "ms": round(ms, 3),
                          "gbs": round(bpe * n / ms / 1e6, 1), "frac_of_hbm_peak": round(bpe * n / ms / 1e6 / PEAK, 3), "bit_exact": ok}), flush=True)
        bi.destroy(); bo.destroy(); sc.destroy(); del out, ref
    del x; torch.cuda.empty_cache()
    xf = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    sc = clo.CloScan("blelloch", ctx, clo.FLOAT, clo.FLOAT)
    bi, bo = clo.Buffer.wrap_tensor(ctx, xf), clo.Buffer.wrap_tensor(ctx, out)
    ms = timed(lambda: sc.with_device_data(q, bi, bo, n))
    ref = torch.cumsum(xf.to(torch.float64), 0) - xf.to(torch.float64)
    err = (out.to(torch.float64) - ref).abs()
    tol = 1e-5 * ref.abs() + 1e-3
    print(json.dumps({"config": "C4 exclusive scan 2^30 f32 -> f32, 1 GPU", "ms": round(ms, 3), "gbs": round(8.0 * n / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(8.0 * n / ms / 1e6 / PEAK, 3), "tolerance": "|gpu-ref| <= 1e-5*|ref| + 1e-3 vs f64 prefix sum",
                      "within_tolerance": bool((err <= tol).all().item()), "max_rel_err": float((err / ref.abs().clamp_min(1.0)).max().item())}), flush=True)
    bi.destroy(); bo.destroy(); sc.destroy(); del xf, out, ref, err, tol
    torch.cuda.empty_cache()

if "c5" in what:
    G, runs = 1 << 22, 1 << 10
    out = torch.empty(G * runs, dtype=torch.int32, device="cuda")        # 16 GiB
    bo = clo.Buffer.wrap_tensor(ctx, out)
    for name in ("xorshift128", "mwc64x"):
        def gen():
            r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, G, 0, "KNUTH(x)", q)
            r.generate(q, bo, runs)
            return r
        r = gen(); torch.cuda.synchronize(); r.destroy()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, G, 0, "KNUTH(x)", q); torch.cuda.synchronize()
        a.record(); r.generate(q, bo, runs); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b); r.destroy()
        # slice: streams [0, 2^20) x runs [0, 16) and the last 16 runs, bit-exact vs the oracle
        SG, SR = 1 << 20, 16
        seeds = oracle.rng_seeds_dev_gid(name, 1, 0, SG)
        want, _ = oracle.rng_generate(name, seeds, SG, SR)
        got = out.view(runs, G)[:SR, :SG].contiguous().cpu().numpy().view(np.uint32).reshape(-1)
        ok = bool(np.array_equal(got, np.asarray(want).reshape(-1)))
        print(json.dumps({"config": "C5 bulk generation 2^32 words (2^22 streams x 2^10 runs), 1 GPU", "rng": name, "ms": round(ms, 3),
                          "gwords_per_s": round(G * runs / ms / 1e6, 1), "gbs": round(4.0 * G * runs / ms / 1e6, 1),
                          "frac_of_hbm_peak": round(4.0 * G * runs / ms / 1e6 / PEAK, 3), "slice_bit_exact_vs_oracle": ok,
                          "slice": "streams [0,2^20) x runs [0,16)"}), flush=True)
    bo.destroy()
