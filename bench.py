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
N * 2^28 keys as one sequence with the sample sort of cl_ops_b200/dist.py (NCCL
all-to-all-v exchange); value = all keys / max-over-ranks time.

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
    sample = "2^%d uint32 keys per step (bounded sample of the 2^28 workload), oracle port of satradix radix=16, OpenMP" % log2n
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
        cnt = torch.tensor([u.numel()], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        e = edges.tolist()
        ok = ok and all(e[2 * i + 1] <= e[2 * i + 2] for i in range(world - 1)) and int(cnt.item()) == world * n
        okt = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(okt.item())
        dbg = ops.sorter.debug(queue)
        del u, k_sorted

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timing
    e2e = None
    e2e_steps = max(1, min(3, args.steps))
    h_in = torch.empty(n, dtype=torch.int32).pin_memory()
    h_in.copy_(t_in.cpu())
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
               "api": "clo_sort_with_host_data (pinned host buffers, device alloc + H2D + sort + D2H per call)"}
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
                       + (" (sample sort across %d GPUs, NCCL all-to-all-v)" % world if distributed else ""),
                       "keys_per_gpu": n, "keys": "xorshift128 DEV_GID KNUTH(x) main_seed 0",
                       "api": "clo_sort_new('satradix') + clo_sort_with_device_data (out of place)",
                       "l2": "inputs (1 GiB/GPU) larger than L2, no flush", "parallelism": "sample-sort x%d" % world},
            "clocks": clocks, "gpu_launches": int(launches), "verified_sorted": ok,
            "repaired_tiles": int(dbg[1]), "lookback_timeout": int(dbg[0]),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
        }
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
