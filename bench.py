#!/usr/bin/env python3
"""bench.py -- the headline measurement of the cl_ops hot path on B200.

Workload (BASELINE.json configs[1]): satradix-equivalent LSD radix sort of 2^28 uint32
keys through the reference's own entry points (clo_sort_new("satradix") +
clo_sort_with_device_data / clo_sort_with_host_data).  One "step" = one sort of the whole
key vector.  Keys are synthetic: this repo's xorshift128 generator, DEV_GID seeds,
KNUTH(x) hash, main_seed 0 (bit-exact against the oracle, so the input is reproducible).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA)
  python bench.py --impl reference [...]                        the reference's algorithm on
                                                                the host cores (oracle port)

N > 1 (under torchrun): every rank holds 2^28 keys (weak scaling); the ranks sort the
N * 2^28 keys as one sequence with the sample sort of cl_ops_b200/dist.py (fused partition +
peer-memory scatter over NVLink; samples, bucket sizes and the barrier are peer-memory writes with epoch flags,
clo_dist_sort_with_device_data in csrc/dist.cu);
value = all keys / max-over-ranks time.  "secondary" in the JSON line holds the other
BASELINE.json configs measured by the same run: scans of 2^30 elements, 2^32 RNG words, the
2^20 sbitonic sort, the 2^30 key-value sort (uniform and Zipf) and the headline sort on
skewed keys -- each with its parity check.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream, barrier +
synchronize on both sides, max over ranks; inputs (1 GiB per rank) are larger than the
126 MB L2, so no flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2N = 28
METRIC = "sort Gkeys/s (2^28 u32 LSD radix sort)"
UNIT = "Gkeys/s"
BYTES_PER_KEY = 36.0          # 4 B histogram read + 4 passes x (4 B read + 4 B write)
BYTES_PER_KEY_PASS = 8.0      # one onesweep launch: read + write of every key
FALLBACK_PEAK = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_PEAK, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's algorithm restated for the host cores
# --------------------------------------------------------------------------------------

def cpu_sort_gkeys(log2n, steps, warmup, threads):
    """satradix port (radix 16, 8 passes of {tile sort, histogram, scan, scatter}) on the host."""
    import oracle
    n = 1 << log2n
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    for _ in range(warmup):
        oracle.sort_satradix(keys, oracle.UINT, radix=16, lws=256, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = oracle.sort_satradix(keys, oracle.UINT, radix=16, lws=256, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    assert bool(np.all(out[1:] >= out[:-1]))
    return n / dt / 1e9, dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    log2n = 24
    value, ms = cpu_sort_gkeys(log2n, max(1, args.steps), min(args.warmup, 1), threads)
    sample = ("2^%d uint32 keys per step (a bounded sample of the 2^28 workload: NOT the same size as the GPU arm), oracle port of "
              "satradix radix=16 -- 8 passes of tile sort + histogram + scan + scatter, the tile sort done as a counting sort "
              "instead of the reference's 1-bit-split scan rounds (same result, cheaper on a CPU) -- OpenMP on all host cores" % log2n)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": "satradix-equivalent LSD radix sort of uint32 keys on the host cores",
                   "log2_keys_per_step": log2n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------
# secondary lines: the other BASELINE.json configs, measured by the same run
# --------------------------------------------------------------------------------------

def _dev_ms(torch, fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def _xor_reduce(torch, x):
    """xor of all elements of an int64 tensor (torch has no xor reduction: halve until one is left)."""
    acc = 0
    while x.numel() > 1:
        if x.numel() % 2:
            acc ^= int(x[-1].item())
            x = x[:-1]
        h = x.numel() // 2
        x = torch.bitwise_xor(x[:h], x[h:])
    return acc ^ (int(x[0].item()) if x.numel() else 0)


def _zipf_u32(torch, n, seed):
    """Zipf(s = 1) over 2^20 distinct values, ranks hashed with Knuth's multiplier."""
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    u = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
    rank = torch.exp(u * float(np.log(1 << 20))).to(torch.int64).clamp_(1, 1 << 20)
    return ((rank * 2654435761) & 0xFFFFFFFF).to(torch.int32)


def _zipf_u64(torch, n, seed):
    """Zipf(s = 1) over 2^24 distinct values, key = rank * 0x9E3779B97F4A7C15 mod 2^64."""
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    u = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
    rank = torch.exp(u * float(np.log(1 << 24))).to(torch.int64).clamp_(1, 1 << 24)
    return rank * -7046029254386353131


def secondary_single(torch, clo, ctx, queue, peak, sorter, log2n):
    """configs C1, C2 (skewed), C3, C4, C5 on one GPU: device time (CUDA events), inputs in HBM."""
    W = clo.Buffer.wrap_tensor
    sec = {}
    n = 1 << log2n
    # -- C2 on skewed keys: Zipf(1.0) and already sorted input, the headline sorter
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    bo = W(ctx, out)
    for name, keys in (("sort_u32_zipf", _zipf_u32(torch, n, 1)), ("sort_u32_sorted", torch.arange(n, dtype=torch.int32, device="cuda"))):
        bi = W(ctx, keys)
        ms = _dev_ms(torch, lambda: sorter.with_device_data(queue, bi, bo, n))
        u = out.to(torch.int64) & 0xFFFFFFFF
        ok = bool((u[1:] >= u[:-1]).all().item()) and int(u.sum().item()) == int((keys.to(torch.int64) & 0xFFFFFFFF).sum().item())
        sec[name] = {"n": n, "ms": ms, "gkeys_s": n / ms / 1e6, "frac": BYTES_PER_KEY * n / ms / 1e6 / peak, "sorted_permutation": ok}
        bi.destroy(); del keys, u
    bo.destroy(); del out
    torch.cuda.empty_cache()
    # -- C1: sbitonic, 2^20 u32 (the reference's CPU-runnable case)
    nb = 1 << 20
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    a0 = torch.randint(-2**31, 2**31 - 1, (nb,), dtype=torch.int32, device="cuda", generator=g)
    t = torch.empty_like(a0)
    bt = W(ctx, t)
    sb = clo.CloSort("sbitonic", ctx, clo.UINT)

    def run_bitonic():
        t.copy_(a0)
        sb.with_device_data(queue, bt, None, nb)
    ms_all = _dev_ms(torch, run_bitonic, iters=20)
    ms_copy = _dev_ms(torch, lambda: t.copy_(a0), iters=20)
    run_bitonic()
    u = t.to(torch.int64) & 0xFFFFFFFF
    ok = bool((u[1:] >= u[:-1]).all().item()) and int(u.sum().item()) == int((a0.to(torch.int64) & 0xFFFFFFFF).sum().item())
    sec["sbitonic_u32_2p20"] = {"n": nb, "ms": ms_all - ms_copy, "mkeys_s": nb / (ms_all - ms_copy) / 1e3, "sorted_permutation": ok,
                                "note": "in place; the restoring copy of the input is timed separately and subtracted"}
    bt.destroy(); sb.destroy(); del a0, t, u
    # -- C4: exclusive scans of 2^30 elements
    ns = 1 << 30
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    x = torch.randint(0, 128, (ns,), dtype=torch.int32, device="cuda", generator=g)
    for name, st, sdt, bpe in (("scan_u32_u32", clo.UINT, torch.int32, 8), ("scan_u32_u64", clo.ULONG, torch.int64, 12)):
        o = torch.empty(ns, dtype=sdt, device="cuda")
        sc = clo.CloScan("blelloch", ctx, clo.UINT, st)
        bi, bo = W(ctx, x), W(ctx, o)
        ms = _dev_ms(torch, lambda: sc.with_device_data(queue, bi, bo, ns))
        ref = torch.cumsum(x.to(torch.int64), 0) - x
        m = 0xFFFFFFFF if sdt == torch.int32 else -1
        ok = bool(torch.equal(o.to(torch.int64) & m, ref & m))
        sec[name] = {"n": ns, "ms": ms, "gbs": bpe * ns / ms / 1e6, "frac": bpe * ns / ms / 1e6 / peak, "bytes_per_elem": bpe, "bit_exact": ok}
        bi.destroy(); bo.destroy(); sc.destroy(); del o, ref
    # -- C4 end to end: clo_scan_with_host_data on pinned host buffers, 2^28 elements (1 GiB each way).
    #    The call pipelines 2^24-element chunks (copy in | scan with carry | copy out), so both directions
    #    of the host link are busy at once; beside it the same bytes as plain sequential pinned copies.
    ne = 1 << 28
    h_x = torch.empty(ne, dtype=torch.int32).pin_memory(); h_x.copy_(x[:ne].cpu())
    h_o = torch.empty(ne, dtype=torch.int32).pin_memory()
    d_tmp = torch.empty(ne, dtype=torch.int32, device="cuda")
    sc = clo.CloScan("blelloch", ctx, clo.UINT, clo.UINT)
    sc.with_host_pointers(h_x.data_ptr(), h_o.data_ptr(), ne, queue)
    t0 = time.perf_counter()
    for _ in range(5):
        sc.with_host_pointers(h_x.data_ptr(), h_o.data_ptr(), ne, queue)
    dt = (time.perf_counter() - t0) / 5
    ref = (torch.cumsum(x[:ne].to(torch.int64), 0) - x[:ne]) & 0xFFFFFFFF
    ok = bool(torch.equal(h_o.to(torch.int64).cuda() & 0xFFFFFFFF, ref))
    d_tmp.copy_(h_x, non_blocking=True); h_o.copy_(d_tmp, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d_tmp.copy_(h_x, non_blocking=True); h_o.copy_(d_tmp, non_blocking=True); torch.cuda.synchronize()
    fl = (time.perf_counter() - t0) / 5
    sec["scan_u32_u32_e2e"] = {"n": ne, "ms": dt * 1e3, "gbs": 8.0 * ne / dt / 1e9, "h2d_bytes": 4 * ne, "d2h_bytes": 4 * ne,
                               "sequential_copies_ms": fl * 1e3, "speedup_vs_sequential_copies": fl / dt, "bit_exact": ok,
                               "api": "clo_scan_with_host_data (pinned host buffers; chunks of 2^24 elements pipelined: H2D | scan | D2H)"}
    sc.destroy(); del h_x, h_o, d_tmp, ref
    del x
    torch.cuda.empty_cache()
    xf = torch.rand(ns, dtype=torch.float32, device="cuda", generator=g)
    o = torch.empty(ns, dtype=torch.float32, device="cuda")
    sc = clo.CloScan("blelloch", ctx, clo.FLOAT, clo.FLOAT)
    bi, bo = W(ctx, xf), W(ctx, o)
    ms = _dev_ms(torch, lambda: sc.with_device_data(queue, bi, bo, ns))
    ref = torch.cumsum(xf.to(torch.float64), 0) - xf.to(torch.float64)
    err = (o.to(torch.float64) - ref).abs()
    sec["scan_f32"] = {"n": ns, "ms": ms, "gbs": 8.0 * ns / ms / 1e6, "frac": 8.0 * ns / ms / 1e6 / peak, "bytes_per_elem": 8,
                       "tolerance": "|gpu - ref| <= 1e-5 * |ref| + 1e-3 against an f64 prefix sum",
                       "within_tolerance": bool((err <= 1e-5 * ref.abs() + 1e-3).all().item()),
                       "max_rel_err": float((err / ref.abs().clamp_min(1.0)).max().item())}
    bi.destroy(); bo.destroy(); sc.destroy(); del xf, o, ref, err
    torch.cuda.empty_cache()
    # -- C5: 2^32 words of xorshift128 / mwc64x (2^22 streams x 2^10 runs), slice checked against the oracle
    import oracle
    G, runs = 1 << 22, 1 << 10
    o = torch.empty(G * runs, dtype=torch.int32, device="cuda")
    bo = W(ctx, o)
    for name in ("xorshift128", "mwc64x"):
        r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, G, 0, "KNUTH(x)", queue)
        r.generate(queue, bo, runs); torch.cuda.synchronize(); r.destroy()
        r = clo.CloRng(name, ctx, clo.SEED_DEV_GID, None, G, 0, "KNUTH(x)", queue)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r.generate(queue, bo, runs); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b); r.destroy()
        SG, SR = 1 << 16, 16
        seeds = oracle.rng_seeds_dev_gid(name, 1, 0, SG)
        want, _ = oracle.rng_generate(name, seeds, SG, SR)
        got = o.view(runs, G)[:SR, :SG].contiguous().cpu().numpy().view(np.uint32).reshape(-1)
        sec["rng_" + name] = {"words": G * runs, "ms": ms, "gwords_s": G * runs / ms / 1e6, "gbs": 4.0 * G * runs / ms / 1e6,
                              "frac": 4.0 * G * runs / ms / 1e6 / peak, "slice": "streams [0, 2^16) x runs [0, 16)",
                              "slice_bit_exact_vs_oracle": bool(np.array_equal(got, np.asarray(want).reshape(-1)))}
    bo.destroy(); del o
    torch.cuda.empty_cache()
    # -- C3: key-value sort of 2^30 (u64 key, u32 payload = index), uniform and Zipf
    nk = 1 << 30
    s64 = clo.CloSort("satradix", ctx, clo.ULONG)
    for name in ("kv_u64_u32_uniform", "kv_u64_u32_zipf"):
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        keys0 = torch.randint(-2**63, 2**63 - 1, (nk,), dtype=torch.int64, device="cuda", generator=g) if name.endswith("uniform") else _zipf_u64(torch, nk, 3)
        keys = keys0.clone(); pay = torch.arange(nk, dtype=torch.int32, device="cuda")
        bk, bp = W(ctx, keys), W(ctx, pay)
        s64.pairs_with_device_data(queue, bk, bp, nk); torch.cuda.synchronize()
        keys.copy_(keys0); pay.copy_(torch.arange(nk, dtype=torch.int32, device="cuda")); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s64.pairs_with_device_data(queue, bk, bp, nk); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        uk = keys ^ (-2**63)
        srt = bool((uk[1:] >= uk[:-1]).all().item())
        eq = uk[1:] == uk[:-1]
        stable = bool((pay[1:][eq] > pay[:-1][eq]).all().item())
        del uk, eq
        src_ok = bool(torch.equal(keys0[pay.to(torch.int64)], keys))
        sec[name] = {"n": nk, "ms": ms, "gpairs_s": nk / ms / 1e6, "gbs": 200.0 * nk / ms / 1e6, "frac": 200.0 * nk / ms / 1e6 / peak,
                     "bytes_per_pair": 200, "sorted": srt, "stable": stable, "payload_is_source_index": src_ok}
        bk.destroy(); bp.destroy(); del keys, pay, keys0
        torch.cuda.empty_cache()
    s64.destroy()
    return sec


def secondary_multi(torch, dist, clo, cdist, ctx, queue, peak, rank, world):
    """configs C3, C4, C5 across the GPUs of the job: device time, max over ranks."""
    W = clo.Buffer.wrap_tensor
    sec = {}

    def maxms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def all_ok(ok):
        t = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN); return bool(t.item())
    # -- C4: 2^30 elements per GPU: reduce -> all-gather of totals -> scan with a device-resident carry
    n = 1 << 30
    for name, et, st, edt in (("scan_u32_u32", clo.UINT, clo.UINT, torch.int32), ("scan_f32", clo.FLOAT, clo.FLOAT, torch.float32)):
        g = torch.Generator(device="cuda"); g.manual_seed(7 + rank)
        x = torch.randint(0, 128, (n,), dtype=torch.int32, device="cuda", generator=g) if edt == torch.int32 else torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
        out = torch.empty(n, dtype=edt, device="cuda")
        sc = clo.CloScan("blelloch", ctx, et, st)
        tot = torch.zeros(1, dtype=edt, device="cuda")

        def red(d):
            b1, b2 = W(ctx, d), W(ctx, tot)
            sc.reduce_with_device_data(queue, b1, b2, d.numel()); b1.destroy(); b2.destroy()
            return tot

        def scan(d, carry):
            b1, b2, b3 = W(ctx, d), W(ctx, out), W(ctx, carry)
            sc.with_device_data(queue, b1, b2, d.numel(), carry_in=b3); b1.destroy(); b2.destroy(); b3.destroy()
            return out
        for _ in range(2):
            cdist.dist_scan(red, scan, x, edt)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            cdist.dist_scan(red, scan, x, edt)
        b.record(); torch.cuda.synchronize()
        ms = maxms(a.elapsed_time(b) / 5)
        wide = torch.int64 if edt == torch.int32 else torch.float64
        mytot = x.to(wide).sum().reshape(1)
        alltot = torch.empty(world, dtype=wide, device="cuda"); dist.all_gather_into_tensor(alltot, mytot)
        carry = alltot[:rank].sum()
        ref = torch.cumsum(x.to(wide), 0) - x.to(wide) + carry
        if edt == torch.int32:
            ok = all_ok(bool(torch.equal(out.to(torch.int64) & 0xFFFFFFFF, ref & 0xFFFFFFFF)))
            extra = {"bit_exact": ok}
        else:
            err = (out.to(torch.float64) - ref).abs()
            ok = all_ok(bool((err <= 1e-5 * ref.abs() + 1e-3).all().item()))
            extra = {"within_tolerance": ok, "tolerance": "|gpu - ref| <= 1e-5 * |ref| + 1e-3 against an f64 prefix sum"}
            del err
        sec[name] = {"n_per_gpu": n, "ms": ms, "aggregate_gbs": world * 12.0 * n / ms / 1e6, "bytes_per_elem": 12,
                     "frac_per_gpu": 12.0 * n / ms / 1e6 / peak}
        sec[name].update(extra)
        sc.destroy(); del x, out, ref
        torch.cuda.empty_cache()
    # -- C5: 2^32 words over the GPUs, streams partitioned by a gid offset (no communication)
    import oracle
    G_total, runs = 1 << 22, 1 << 10
    first, count = cdist.rng_partition(G_total)
    o = torch.empty(count * runs, dtype=torch.int32, device="cuda")
    bo = W(ctx, o)
    for name in ("xorshift128", "mwc64x"):
        rg = clo.CloRng(name, ctx, seeds_count=count, main_seed=0, hash="KNUTH(x)", queue=queue, gid_offset=first)
        rg.generate(queue, bo, runs); torch.cuda.synchronize(); rg.destroy()
        rg = clo.CloRng(name, ctx, seeds_count=count, main_seed=0, hash="KNUTH(x)", queue=queue, gid_offset=first)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rg.generate(queue, bo, runs); b.record(); torch.cuda.synchronize()
        ms = maxms(a.elapsed_time(b)); rg.destroy()
        SG, SR = 4096, 8
        ok = True
        for g0 in (0, count - SG):
            seeds = oracle.rng_seeds_dev_gid(name, 1, 0, SG, gid0=first + g0)
            want, _ = oracle.rng_generate(name, seeds, SG, SR)
            got = o.view(runs, count)[:SR, g0:g0 + SG].contiguous().cpu().numpy().view(np.uint32)
            ok &= bool(np.array_equal(got, want))
        sec["rng_" + name] = {"words": G_total * runs, "ms": ms, "aggregate_gwords_s": G_total * runs / ms / 1e6,
                              "aggregate_gbs": 4.0 * G_total * runs / ms / 1e6, "slices_bit_exact_vs_oracle": all_ok(ok)}
    bo.destroy(); del o
    torch.cuda.empty_cache()
    # -- C3: key-value sample sort of 2^30 pairs in total (u64 key, u32 payload = global index)
    n = (1 << 30) // world
    ops = cdist.GpuOps(clo, ctx, queue, clo.ULONG)
    ops.setup_peer_exchange(n + n // 2, torch.int64, True)
    for name in ("kv_u64_u32_uniform", "kv_u64_u32_zipf"):
        g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
        keys = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g) if name.endswith("uniform") else _zipf_u64(torch, n, 100 + rank)
        pay = (torch.arange(n, dtype=torch.int64, device="cuda") + rank * n).to(torch.int32)
        for _ in range(2):
            cdist.sample_sort(keys, pay, ops, 64, gidx0=rank * n)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            k, pp, info = cdist.sample_sort(keys, pay, ops, 64, gidx0=rank * n)
        b.record(); torch.cuda.synchronize()
        ms = maxms(a.elapsed_time(b) / 3)
        uk = k ^ (-2**63)
        ok = bool((uk[1:] >= uk[:-1]).all().item()) if uk.numel() > 1 else True
        eq = uk[1:] == uk[:-1]
        pl = pp.to(torch.int64) & 0xFFFFFFFF
        stable = bool((pl[1:][eq] > pl[:-1][eq]).all().item())      # payload = global index (2^30 pairs fit 32 bits)
        edges = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
        if uk.numel():
            edges[2 * rank], edges[2 * rank + 1] = uk[0], uk[-1]
        dist.all_reduce(edges)
        # permutation: count, sum and xor of keys and of payloads over all ranks, before and after
        chk = torch.stack([torch.tensor(k.numel(), device="cuda"), k.sum(), pl.sum(), -torch.tensor(n, device="cuda"), -keys.sum(), -(pay.to(torch.int64) & 0xFFFFFFFF).sum()]).to(torch.int64)
        dist.all_reduce(chk)
        e = edges.tolist()
        c = chk.tolist()
        ok = ok and all(e[2 * i + 1] <= e[2 * i + 2] for i in range(world - 1)) and c[0] + c[3] == 0 and c[1] + c[4] == 0 and c[2] + c[5] == 0
        sec[name] = {"pairs_total": world * n, "ms": ms, "gpairs_s": world * n / ms / 1e6, "sorted_permutation": all_ok(ok),
                     "globally_stable": all_ok(stable), "fused_peer_scatter": bool(info.get("fused")), "received_rank0": info["received"]}
        del keys, pay, k, pp, uk, eq, pl
        torch.cuda.empty_cache()
    ops.close()
    return sec


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import cl_ops_b200 as clo
    from cl_ops_b200 import dist as cdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the host arm")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    clo.lib()
    peak, peak_src = measured_peak()
    n = 1 << args.log2n
    ctx = clo.Context(local_rank)
    queue = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)

    # ---- synthetic keys, resident in HBM: xorshift128 / DEV_GID / KNUTH / seed 0, this rank's streams
    t_in = torch.empty(n, dtype=torch.int32, device="cuda")
    b_in = clo.Buffer.wrap_tensor(ctx, t_in)
    chunk = 1 << 24                      # 2^24 streams x 16 B of state at a time
    for c0 in range(0, n, chunk):
        cnt = min(chunk, n - c0)
        r = clo.CloRng("xorshift128", ctx, seeds_count=cnt, main_seed=0, hash="KNUTH(x)", queue=queue,
                       gid_offset=rank * n + c0)
        sub = clo.Buffer(ctx, size=cnt * 4, ptr=t_in.data_ptr() + c0 * 4)
        r.generate(queue, sub, 1)
        queue.finish()
        sub.destroy()
        r.destroy()
    t_out = torch.empty_like(t_in)
    b_out = clo.Buffer.wrap_tensor(ctx, t_out)
    sorter = clo.CloSort("satradix", ctx, clo.UINT)
    ops = cdist.GpuOps(clo, ctx, queue, clo.UINT) if distributed else None
    if distributed and os.environ.get("CLO_DIST_EXCHANGE", "fused") == "fused":
        ops.setup_peer_exchange(n + n // 4, torch.int32, False)

    def step():
        if distributed:
            return cdist.sample_sort(t_in, None, ops, 32, gidx0=rank * n)
        sorter.with_device_data(queue, b_in, b_out, n)
        return None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, device time, per-kernel events for the roofline
    sorter.set_timing(True)
    if distributed:
        ops.sorter.set_timing(True)
    launches0 = clo.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    pass_ms, hist_ms = [], []
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
        # reading the per-kernel events blocks on the step that just ran; the next step's
        # launch gap is outside every kernel's own duration
        tm = (ops.sorter if distributed else sorter).get_timing()
        if len(tm) >= 2:
            hist_ms.append(tm[0])
            pass_ms.extend(tm[1:])
    barrier()
    clocks = sampler.stop()
    launches = clo.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[args.steps])
    sorter.set_timing(False)
    if distributed:
        ops.sorter.set_timing(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / ms_per_step / 1e6

    # ---- correctness of what was timed
    if not distributed:
        u = t_out.to(torch.int64) & 0xFFFFFFFF
        ok = bool((u[1:] >= u[:-1]).all().item()) and int(u.sum().item()) == int((t_in.to(torch.int64) & 0xFFFFFFFF).sum().item())
        del u
        dbg = sorter.debug(queue)
    else:
        k_sorted, _, info = cdist.sample_sort(t_in, None, ops, 32, profile=True, gidx0=rank * n)
        phases = info.get("phases_ms", {})
        u = k_sorted.to(torch.int64) & 0xFFFFFFFF
        ok = bool((u[1:] >= u[:-1]).all().item()) if u.numel() > 1 else True
        edges = torch.zeros(2 * world, dtype=torch.int64, device="cuda")
        if u.numel():
            edges[2 * rank], edges[2 * rank + 1] = u[0], u[-1]
        dist.all_reduce(edges)
        # a permutation of the input: count, sum and xor of all keys, before and after
        uin = t_in.to(torch.int64) & 0xFFFFFFFF
        xin, xout = _xor_reduce(torch, uin), _xor_reduce(torch, u)
        cnt = torch.stack([torch.tensor(u.numel(), device="cuda") - n, u.sum() - uin.sum()]).to(torch.int64)
        dist.all_reduce(cnt)
        xr = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(world)]
        mine = torch.tensor([xin, xout], dtype=torch.int64, device="cuda")
        dist.all_gather(xr, mine)
        x_in = x_out = 0
        for v in xr:
            a_, b_ = v.tolist()
            x_in ^= a_; x_out ^= b_
        del uin
        e = edges.tolist()
        ok = ok and all(e[2 * i + 1] <= e[2 * i + 2] for i in range(world - 1)) and cnt.tolist() == [0, 0] and x_in == x_out
        okt = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
        dbg = ops.sorter.debug(queue)
        del u, k_sorted

    # ---- the other BASELINE.json configs, same run (device time)
    secondary = None
    if not args.no_secondary:
        if distributed:
            secondary = secondary_multi(torch, dist, clo, cdist, ctx, queue, peak, rank, world)
        else:
            del t_out
            b_out.destroy()
            torch.cuda.empty_cache()
            secondary = secondary_single(torch, clo, ctx, queue, peak, sorter, args.log2n)
            t_out = torch.empty_like(t_in)
            b_out = clo.Buffer.wrap_tensor(ctx, t_out)

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timing
    e2e = None
    e2e_steps = max(10, min(20, args.steps))
    h_in = torch.empty(n, dtype=torch.int32).pin_memory()
    h_in.copy_(t_in.cpu())
    # the floor of any host-buffer path on this box: the same bytes in and out as plain pinned
    # copies, one cudaMemcpyAsync each way, nothing else
    h_floor = torch.empty(n, dtype=torch.int32).pin_memory()
    d_floor = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        d_floor.copy_(h_in, non_blocking=True); h_floor.copy_(d_floor, non_blocking=True); torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        d_floor.copy_(h_in, non_blocking=True); h_floor.copy_(d_floor, non_blocking=True); torch.cuda.synchronize()
    floor_dt = (time.perf_counter() - t0) / 5
    if distributed:
        tf = torch.tensor([floor_dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        floor_dt = float(tf.item())
    del h_floor, d_floor
    if not distributed:
        h_out = torch.empty(n, dtype=torch.int32).pin_memory()
        e2e_sorter = clo.CloSort("satradix", ctx, clo.UINT)

        def e2e_step():
            e2e_sorter.with_host_pointers(h_in.data_ptr(), h_out.data_ptr(), n, queue)
        e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()          # blocks until the sorted keys are back in host memory
        dt = (time.perf_counter() - t0) / e2e_steps
        ho = h_out.numpy().view(np.uint32)
        assert bool(np.all(ho[1:] >= ho[:-1]))
        e2e = {"value": n / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": 4 * n,
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "n_gpus": 1,
               "pcie_floor_ms": floor_dt * 1e3, "frac_of_pcie_floor": floor_dt / dt,
               "api": "clo_sort_with_host_data (pinned host buffers: H2D + sort + D2H per call)"}
        e2e_sorter.destroy()
        del h_out
    else:
        # every rank: its shard host -> device, the sample sort, its slice of the result device -> host
        h_out = torch.empty(n + n // 4, dtype=torch.int32).pin_memory()
        d_in = torch.empty(n, dtype=torch.int32, device="cuda")

        def e2e_step():
            d_in.copy_(h_in, non_blocking=True)
            k, _, _ = cdist.sample_sort(d_in, None, ops, 32, gidx0=rank * n)
            h_out[:k.numel()].copy_(k, non_blocking=True)
            torch.cuda.synchronize()
            return k.numel()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        got = 0
        for _ in range(e2e_steps):
            got = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        ho = h_out.numpy().view(np.uint32)[:got]
        assert bool(np.all(ho[1:] >= ho[:-1]))
        e2e = {"value": world * n / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 4 * n * world,
               "d2h_bytes_per_step": 4 * n * world, "ms_per_step": dt * 1e3, "steps": e2e_steps, "n_gpus": world,
               "pcie_floor_ms": floor_dt * 1e3, "frac_of_pcie_floor": floor_dt / dt,
               "api": "per rank: pinned host shard -> device, cl_ops_b200.dist.sample_sort, sorted slice -> pinned host; max over ranks"}
        del h_out, d_in
    del h_in

    # ---- CPU baseline beside it (rank 0, N == 1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, ms = cpu_sort_gkeys(24, 2, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 sorts of 2^24 uint32 keys (%.0f ms each), oracle port of the reference's satradix "
                         "(radix 16: 8 passes of tile sort + histogram + scan + scatter), OpenMP on all host cores" % ms}

    if rank == 0:
        roof = None
        if pass_ms:
            avg = float(np.mean(pass_ms))
            ach = BYTES_PER_KEY_PASS * n / avg / 1e6
            roof = {"bound": "hbm", "kernel": "clo_radix_onesweep_v6", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": BYTES_PER_KEY_PASS * n, "avg_launch_ms": avg,
                    "launches_timed": len(pass_ms), "histogram_ms": float(np.mean(hist_ms)),
                    "share_of_step": 4 * avg / ms_per_step,
                    "note": "per GPU: the onesweep passes of the local sort (rank 0)" if distributed else None,
                    "whole_sort": {"achieved": BYTES_PER_KEY * n / ms_per_step / 1e6,
                                   "frac": BYTES_PER_KEY * n / ms_per_step / 1e6 / peak, "bytes_per_key": BYTES_PER_KEY}}
            try:
                with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                    roof["traffic"] = json.load(f).get("clo_radix_onesweep_bytes_per_launch_2^28")
            except Exception:
                pass
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "satradix-equivalent LSD radix sort of 2^%d uint32 keys per GPU" % args.log2n
                       + (" (sample sort across %d GPUs: clo_dist_sort_with_device_data: fused partition + CUDA-IPC peer-memory scatter over NVLink; samples, sizes and the barrier are peer-memory writes with epoch flags)" % world if distributed else ""),
                       "keys_per_gpu": n, "keys": "xorshift128 DEV_GID KNUTH(x) main_seed 0",
                       "api": "clo_sort_new('satradix') + clo_sort_with_device_data (out of place)",
                       "l2": "inputs (1 GiB/GPU) larger than L2, no flush", "parallelism": "sample-sort x%d" % world},
            "clocks": clocks, "gpu_launches": int(launches), "verified_sorted": ok,
            "repaired_tiles": int(dbg[1]), "lookback_timeout": int(dbg[0]),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "secondary": secondary,
        }
        # the vendor library on the same kind of box, measured once by tools/cub_yardstick.cu (never
        # linked into the product); context for the absolute numbers, not part of the metric
        try:
            with open(os.path.join(ROOT, "profiles", "r2_cub_yardstick.jsonl")) as f:
                ys = [json.loads(x) for x in f if x.strip()]
            out["yardstick"] = {"source": "profiles/r2_cub_yardstick.jsonl (tools/cub_yardstick.cu, one B200, recorded; not run here)",
                                "cub_sort_keys_2p28_u32_gkeys_s": next(y["gkeys_s"] for y in ys if "SortKeys" in y["yardstick"]),
                                "cub_sort_pairs_u64_u32_gpairs_s": next(y["gpairs_s"] for y in ys if "SortPairs" in y["yardstick"]),
                                "cub_exclusive_sum_2p30_u32_gbs": next(y["gbs"] for y in ys if "ExclusiveSum" in y["yardstick"] and y["dtype"] == "u32"),
                                "cub_exclusive_sum_2p30_f32_gbs": next(y["gbs"] for y in ys if "ExclusiveSum" in y["yardstick"] and y["dtype"] == "f32")}
        except Exception:
            pass
        if distributed:
            out["phases_ms_rank0"] = {k: round(v, 3) for k, v in phases.items()}
            sc = phases.get("scatter to peers")
            out["exchange"] = {"sent_keys_rank0": info["sent"], "received_keys_rank0": info["received"],
                               "fused_peer_scatter": bool(info.get("fused"))}
            if sc:
                # NVLink roofline of the exchange: bytes this rank pushes to its peers / the
                # duration of the fused partition-scatter kernel that pushes them
                out["exchange"].update({"remote_bytes_rank0": 4 * info["sent"], "scatter_ms": sc,
                                        "nvlink_gbs": 4 * info["sent"] / sc / 1e6, "nvlink_peak_gbs": 900.0,
                                        "nvlink_frac": 4 * info["sent"] / sc / 1e6 / 900.0})
        print(json.dumps(out))
    b_in.destroy(); b_out.destroy(); sorter.destroy()
    if ops:
        ops.close()
    queue.destroy(); ctx.destroy()
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2N)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary lines (other BASELINE configs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
