"""tools/dist_scan_rng_run.py -- BASELINE.json configs[3] and [4] across N GPUs (torchrun):
  scan  2^30 u32 (and f32) elements PER GPU: per-GPU reduce -> all-gather of N totals -> per-GPU
        single-pass scan with a device-resident carry-in; bit-exact vs a global int64 prefix sum
        (checked through the carry of every rank and the first/last elements of every shard)
  rng   2^32 / N words per GPU of xorshift128 / mwc64x, streams partitioned by a gid offset, no
        communication; rank r's first and last stream slices bit-exact vs the oracle
Device time (CUDA events), max over ranks; one JSON line per case on rank 0."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo
from cl_ops_b200 import dist as cdist
import oracle

r, P, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = clo.Context(lr); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
PEAK = 6542.1
W = clo.Buffer.wrap_tensor


def maxms(ms):
    t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())


n = 1 << int(os.environ.get("LOG2N", "30"))
for et, st, edt, sdt, label, bpe in ((clo.UINT, clo.UINT, torch.int32, torch.int32, "u32 -> u32", 8), (clo.FLOAT, clo.FLOAT, torch.float32, torch.float32, "f32 -> f32", 8)):
    g = torch.Generator(device="cuda"); g.manual_seed(7 + r)
    x = torch.randint(0, 128, (n,), dtype=torch.int32, device="cuda", generator=g) if edt == torch.int32 else torch.rand(n, dtype=torch.float32, device="cuda", generator=g)
    out = torch.empty(n, dtype=sdt, device="cuda")
    sc = clo.CloScan("blelloch", ctx, et, st)
    tot = torch.zeros(1, dtype=sdt, device="cuda")

    def red(d):
        b1, b2 = W(ctx, d), W(ctx, tot)
        sc.reduce_with_device_data(q, b1, b2, d.numel()); b1.destroy(); b2.destroy()
        return tot

    def scan(d, carry):
        b1, b2, b3 = W(ctx, d), W(ctx, out), W(ctx, carry)
        sc.with_device_data(q, b1, b2, d.numel(), carry_in=b3); b1.destroy(); b2.destroy(); b3.destroy()
        return out
    for _ in range(2):
        cdist.dist_scan(red, scan, x, sdt)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        cdist.dist_scan(red, scan, x, sdt)
    b.record(); torch.cuda.synchronize()
    ms = maxms(a.elapsed_time(b) / 5)
    # check: my shard against a local prefix sum + the true carry (sum of lower ranks' totals)
    if edt == torch.int32:
        mytot = x.to(torch.int64).sum().reshape(1)
        alltot = torch.empty(P, dtype=torch.int64, device="cuda"); dist.all_gather_into_tensor(alltot, mytot)
        carry = int(alltot[:r].sum().item())
        ref = (torch.cumsum(x.to(torch.int64), 0) - x + carry) & 0xFFFFFFFF
        ok = bool(torch.equal(out.to(torch.int64) & 0xFFFFFFFF, ref))
        extra = {"bit_exact": ok}
    else:
        mytot = x.to(torch.float64).sum().reshape(1)
        alltot = torch.empty(P, dtype=torch.float64, device="cuda"); dist.all_gather_into_tensor(alltot, mytot)
        carry = float(alltot[:r].sum().item())
        ref = torch.cumsum(x.to(torch.float64), 0) - x.to(torch.float64) + carry
        err = (out.to(torch.float64) - ref).abs()
        ok = bool((err <= 1e-5 * ref.abs() + 1e-3).all().item())
        extra = {"within_tolerance": ok, "tolerance": "|gpu-ref| <= 1e-5*|ref| + 1e-3 vs f64 prefix sum"}
    okt = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if r == 0:
        d = {"config": "C4 exclusive scan, 2^%d %s elements per GPU" % (int(np.log2(n)), label), "n_gpus": P, "ms": round(ms, 3),
             "aggregate_gbs": round(P * 12.0 * n / ms / 1e6, 1), "bytes_per_elem": "12 (reduce read 4 + scan 8)",
             "frac_of_hbm_peak_per_gpu": round(12.0 * n / ms / 1e6 / PEAK, 3)}
        d.update(extra); d[list(extra)[0]] = bool(okt.item())
        print(json.dumps(d), flush=True)
    sc.destroy(); del x, out, ref
    torch.cuda.empty_cache()

G_total, runs = 1 << 22, 1 << 10
first, count = cdist.rng_partition(G_total)
out = torch.empty(count * runs, dtype=torch.int32, device="cuda")
bo = W(ctx, out)
for name in ("xorshift128", "mwc64x"):
    rg = clo.CloRng(name, ctx, seeds_count=count, main_seed=0, hash="KNUTH(x)", queue=q, gid_offset=first)
    rg.generate(q, bo, runs); torch.cuda.synchronize(); rg.destroy()
    rg = clo.CloRng(name, ctx, seeds_count=count, main_seed=0, hash="KNUTH(x)", queue=q, gid_offset=first)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rg.generate(q, bo, runs); b.record(); torch.cuda.synchronize()
    ms = maxms(a.elapsed_time(b)); rg.destroy()
    SG, SR = 4096, 8
    ok = True
    for g0 in (0, count - SG):
        seeds = oracle.rng_seeds_dev_gid(name, 1, 0, SG, gid0=first + g0)
        want, _ = oracle.rng_generate(name, seeds, SG, SR)
        got = out.view(runs, count)[:SR, g0:g0 + SG].contiguous().cpu().numpy().view(np.uint32)
        ok &= bool(np.array_equal(got, want))
    okt = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if r == 0:
        print(json.dumps({"config": "C5 bulk generation, 2^32 words over %d GPUs (2^22 streams x 2^10 runs, streams partitioned)" % P, "rng": name,
                          "n_gpus": P, "ms": round(ms, 3), "aggregate_gwords_per_s": round(G_total * runs / ms / 1e6, 1),
                          "aggregate_gbs": round(4.0 * G_total * runs / ms / 1e6, 1), "slices_bit_exact_vs_oracle": bool(okt.item())}), flush=True)
bo.destroy(); q.destroy(); ctx.destroy()
dist.barrier(); dist.destroy_process_group()
