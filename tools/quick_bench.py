"""tools/quick_bench.py -- device-timed throughput of the three hot kernels at the
BASELINE.json sizes (development aid; bench.py is the contract).  CUDA events on the
stream the library launches on (the torch current stream, wrapped as a CCLQueue)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), float(np.min(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n-sort", type=int, default=28)
    ap.add_argument("--log2n-scan", type=int, default=30)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--what", default="sort,sort64,pairs,scan,scanf,rng,bitonic")
    args = ap.parse_args()
    what = args.what.split(",")
    peak = 6542.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    ctx = clo.Context()
    q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
    res = {}

    if "sort" in what:
        n = 1 << args.log2n_sort
        t_in = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        t_out = torch.empty_like(t_in)
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        s = clo.CloSort("satradix", ctx, clo.UINT)
        med, best = timed(lambda: s.with_device_data(q, b_in, b_out, n), args.iters)
        chk = t_out.view(torch.int32)
        u = (chk.to(torch.int64) & 0xFFFFFFFF)
        ok = bool((u[1:] >= u[:-1]).all().item())
        res["sort_u32"] = dict(n=n, ms=med, ms_best=best, gkeys=n / med / 1e6, gbs=36.0 * n / med / 1e6,
                               frac=36.0 * n / med / 1e6 / peak, sorted=ok)
        print(json.dumps({"sort_u32": res["sort_u32"]}), flush=True)
        b_in.destroy(); b_out.destroy(); s.destroy(); del t_in, t_out, chk, u
        torch.cuda.empty_cache()

    if "sortprof" in what:
        n = 1 << args.log2n_sort
        os.environ["CLO_RADIX_PROFILE"] = "1"
        t_in = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        dist = os.environ.get("QB_DIST", "uniform")
        if dist == "zipf":
            u = torch.rand(n, device="cuda", dtype=torch.float64)
            rank = torch.exp(u * float(np.log(1 << 20))).to(torch.int64).clamp_(1, 1 << 20)
            t_in = ((rank * 2654435761) & 0xFFFFFFFF).to(torch.int32)
            del u, rank
        elif dist == "sorted":
            t_in = torch.arange(n, dtype=torch.int32, device="cuda")
        elif dist == "v16":
            t_in = torch.randint(0, 16, (n,), device="cuda", dtype=torch.int32) * 0x01010101
        t_out = torch.empty_like(t_in)
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        s = clo.CloSort("satradix", ctx, clo.UINT)
        for _ in range(3):
            s.with_device_data(q, b_in, b_out, n)
        d = s.debug(q)
        print("cfg", os.environ.get("CLO_RADIX_CFG"))
        names = ["zero+loadwait", "rank", "digit phase", "stage", "ticket+loadissue", "prefix wait", "write-out"]
        if os.environ.get("CLO_RADIX_KERNEL") is None:
            names = ["P1 count+loadwait", "B1 wait", "P2 digits", "B2 wait", "P5 write-out", "B3 wait", "P3 place+P4 load"]
        if os.environ.get("CLO_RADIX_KERNEL") == "v7":
            names = ["P1 count", "B1 wait", "P2 digits+PREF wait", "B2 wait", "P3 place+load", "B3 wait", "P4 write-out"]
        if os.environ.get("CLO_RADIX_KERNEL") == "classic":
            names = ["ticket+zero", "load", "rank", "digit phase", "stage", "look-back", "write-out"]
        tiles = 4 * ((n + 8191) // 8192)
        tot = sum(d[2:9])
        print(json.dumps({"sortprof": {"timeout": d[0], "repaired_tiles": d[1], "tiles": tiles,
                                       "cycles_per_tile": {k: round(v / tiles, 1) for k, v in zip(names, d[2:9])},
                                       "share": {k: round(v / max(tot, 1), 3) for k, v in zip(names, d[2:9])},
                                       "walk_cycles": round(d[9] / tiles, 1), "walk_rounds": round(d[10] / tiles, 2),
                                       "walk_unpublished": round(d[11] / tiles, 2),
                                       "walk_depth": round(d[12] / tiles, 1),
                                       "prop_cycles_total": d[14], "prop_rounds": d[15],
                                       "prop_idle_rounds": d[16], "prop_idle_cycles": d[17],
                                       "prop_wait_bar_store_cycles": [d[11], d[12], d[13]]}}), flush=True)
        os.environ.pop("CLO_RADIX_PROFILE")
        b_in.destroy(); b_out.destroy(); s.destroy(); del t_in, t_out
        torch.cuda.empty_cache()

    if "partition" in what:
        # the sample sort's stable 8-way partition, local destinations (one GPU): count + offsets + scatter
        n = 1 << args.log2n_sort
        P = 8
        t_in = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        t_out = torch.empty_like(t_in)
        spl = torch.tensor([(i * (1 << 32) // P) - (1 << 32 if i * (1 << 32) // P >= (1 << 31) else 0) for i in range(1, P)], dtype=torch.int32, device="cuda")
        spi = torch.zeros(P - 1, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(P, dtype=torch.int64, device="cuda")
        W = clo.Buffer.wrap_tensor
        b = [W(ctx, t_in), W(ctx, t_out), W(ctx, spl), W(ctx, spi), W(ctx, cnt)]
        s = clo.CloSort("satradix", ctx, clo.UINT)
        med, best = timed(lambda: s.partition_with_device_data(q, b[0], None, b[1], None, n, 0, b[2], b[3], P, b[4]), args.iters)
        print(json.dumps({"partition_u32_8way": dict(n=n, ms=med, gkeys=n / med / 1e6, gbs=12.0 * n / med / 1e6, frac=12.0 * n / med / 1e6 / peak,
                                                    counts=cnt.tolist())}), flush=True)
        for x in b:
            x.destroy()
        s.destroy()

    if "sortcfg" in what:
        # tuning sweep of the headline kernel: tile configuration x match method
        n = 1 << args.log2n_sort
        t_in = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        t_out = torch.empty_like(t_in)
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        for hw in ("atomic",):
            for cfg in [int(x) for x in os.environ.get("QB_CFGS", "0,1,2,3,4").split(",")]:
                os.environ["CLO_RADIX_RANK"] = hw
                os.environ["CLO_RADIX_CFG"] = str(cfg)
                s = clo.CloSort("satradix", ctx, clo.UINT)
                med, best = timed(lambda: s.with_device_data(q, b_in, b_out, n), args.iters)
                u = (t_out.to(torch.int64) & 0xFFFFFFFF)
                ok = bool((u[1:] >= u[:-1]).all().item())
                del u
                print(json.dumps({"sortcfg": dict(hw=hw, cfg=cfg, n=n, ms=med, gkeys=n / med / 1e6,
                                                  frac=36.0 * n / med / 1e6 / peak, sorted=ok)}), flush=True)
                s.destroy()
        os.environ.pop("CLO_RADIX_RANK"); os.environ.pop("CLO_RADIX_CFG")
        b_in.destroy(); b_out.destroy(); del t_in, t_out
        torch.cuda.empty_cache()

    if "sort64" in what:
        n = 1 << (args.log2n_sort - 1)
        t_in = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda")
        t_out = torch.empty_like(t_in)
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        s = clo.CloSort("satradix", ctx, clo.ULONG)
        med, best = timed(lambda: s.with_device_data(q, b_in, b_out, n), args.iters)
        res["sort_u64"] = dict(n=n, ms=med, gkeys=n / med / 1e6, gbs=136.0 * n / med / 1e6,
                               frac=136.0 * n / med / 1e6 / peak)
        print(json.dumps({"sort_u64": res["sort_u64"]}), flush=True)
        b_in.destroy(); b_out.destroy(); s.destroy(); del t_in, t_out
        torch.cuda.empty_cache()

    if "pairs" in what:
        n = 1 << (args.log2n_sort - 1)
        k0 = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda")
        tk = torch.empty_like(k0)
        tp = torch.empty(n, dtype=torch.int32, device="cuda")
        bk, bp = clo.Buffer.wrap_tensor(ctx, tk), clo.Buffer.wrap_tensor(ctx, tp)
        s = clo.CloSort("satradix", ctx, clo.ULONG)

        def run():
            tk.copy_(k0)
            tp.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
            s.pairs_with_device_data(q, bk, bp, n)
        med_all, _ = timed(run, max(3, args.iters // 2))

        def fill():
            tk.copy_(k0)
            tp.copy_(torch.arange(n, dtype=torch.int32, device="cuda"))
        med_fill, _ = timed(fill, max(3, args.iters // 2))
        med = med_all - med_fill
        res["pairs_u64_u32"] = dict(n=n, ms=med, gpairs=n / med / 1e6, gbs=200.0 * n / med / 1e6,
                                    frac=200.0 * n / med / 1e6 / peak)
        print(json.dumps({"pairs_u64_u32": res["pairs_u64_u32"]}), flush=True)
        bk.destroy(); bp.destroy(); s.destroy(); del k0, tk, tp
        torch.cuda.empty_cache()

    for name, et, st, tdt, bytes_per in (("scan", clo.UINT, clo.UINT, torch.int32, 8),
                                         ("scan64", clo.UINT, clo.ULONG, torch.int32, 12),
                                         ("scanf", clo.FLOAT, clo.FLOAT, torch.float32, 8)):
        if name not in what and not (name == "scan64" and "scan" in what):
            continue
        n = 1 << args.log2n_scan
        if tdt == torch.float32:
            t_in = torch.rand(n, dtype=tdt, device="cuda")
        else:
            t_in = torch.randint(0, 128, (n,), dtype=tdt, device="cuda")
        t_out = torch.empty(n, dtype=torch.int64 if st == clo.ULONG else tdt, device="cuda")
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        s = clo.CloScan("blelloch", ctx, et, st)
        med, best = timed(lambda: s.with_device_data(q, b_in, b_out, n), args.iters)
        if tdt != torch.float32:
            # spot check: last element equals the sum of all but the last input
            want = int(t_in[:-1].to(torch.int64).sum().item())
            got = int(t_out[-1].item())
            if st == clo.UINT:
                want &= 0xFFFFFFFF
                got &= 0xFFFFFFFF
            ok = want == got
        else:
            ok = None
        res[name] = dict(n=n, ms=med, ms_best=best, gelem=n / med / 1e6, gbs=bytes_per * n / med / 1e6,
                         frac=bytes_per * n / med / 1e6 / peak, check=ok)
        print(json.dumps({name: res[name]}), flush=True)
        b_in.destroy(); b_out.destroy(); s.destroy(); del t_in, t_out
        torch.cuda.empty_cache()

    if "scancfg" in what:
        n = 1 << args.log2n_scan
        t_in = torch.randint(0, 128, (n,), dtype=torch.int32, device="cuda")
        t_out = torch.empty_like(t_in)
        b_in, b_out = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        for cfg in [int(x) for x in os.environ.get("QB_SCAN_CFGS", "0,1,2,3,4,5").split(",")]:
            os.environ["CLO_SCAN_CFG"] = str(cfg)
            s = clo.CloScan("blelloch", ctx, clo.UINT, clo.UINT)
            med, best = timed(lambda: s.with_device_data(q, b_in, b_out, n), args.iters)
            want = int(t_in[:-1].to(torch.int64).sum().item()) & 0xFFFFFFFF
            got = int(t_out[-1].item()) & 0xFFFFFFFF
            print(json.dumps({"scancfg": dict(cfg=cfg, n=n, ms=med, gbs=8.0 * n / med / 1e6,
                                              frac=8.0 * n / med / 1e6 / peak, check=want == got)}), flush=True)
            s.destroy()
        os.environ.pop("CLO_SCAN_CFG")
        b_in.destroy(); b_out.destroy(); del t_in, t_out
        torch.cuda.empty_cache()

    if "rng" in what:
        G, runs = 1 << 22, 256
        t_out = torch.empty(G * runs, dtype=torch.int32, device="cuda")
        b_out = clo.Buffer.wrap_tensor(ctx, t_out)
        for rng in ("lcg", "xorshift64", "xorshift128", "mwc64x", "parkmiller", "tauslcg"):
            r = clo.CloRng(rng, ctx, clo.SEED_DEV_GID, None, G, 0, "KNUTH(x)", q)
            med, best = timed(lambda: r.generate(q, b_out, runs), max(3, args.iters // 2))
            res["rng_" + rng] = dict(words=G * runs, ms=med, gwords=G * runs / med / 1e6,
                                     gbs=4.0 * G * runs / med / 1e6, frac=4.0 * G * runs / med / 1e6 / peak)
            print(json.dumps({"rng_" + rng: res["rng_" + rng]}), flush=True)
            r.destroy()
        b_out.destroy(); del t_out
        torch.cuda.empty_cache()

    if "bitonic" in what:
        n = 1 << 20
        a0 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
        t = torch.empty_like(a0)
        b = clo.Buffer.wrap_tensor(ctx, t)
        s = clo.CloSort("sbitonic", ctx, clo.UINT)

        def run():
            t.copy_(a0)
            s.with_device_data(q, b, None, n)
        med, best = timed(run, args.iters)
        res["sbitonic_2^20"] = dict(n=n, ms=med, mkeys=n / med / 1e3)
        print(json.dumps({"sbitonic_2^20": res["sbitonic_2^20"]}), flush=True)
        b.destroy(); s.destroy()

    print(json.dumps(res))
    q.destroy(); ctx.destroy()


if __name__ == "__main__":
    main()
