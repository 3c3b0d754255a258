"""tools/dist_config_run.py -- BASELINE.json configs[2] across N GPUs (torchrun): key-value sample
sort of 2^LOG2N (u64 key, u32 payload = low 32 bits of the global index) per GPU, uniform and
Zipf(1.0); fused partition + peer-memory exchange; device time, max over ranks; checks: locally
sorted, rank boundaries ordered, element count, and (uniform case) equal keys keep ascending
payloads inside a rank."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo
from cl_ops_b200 import dist as cdist

r, P, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << int(os.environ.get("LOG2N", "27"))
ctx = clo.Context(lr); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
ops = cdist.GpuOps(clo, ctx, q, clo.ULONG)
ops.setup_peer_exchange(n + n // 2, torch.int64, True)
for name in ("uniform", "zipf(1.0) over 2^24 values"):
    g = torch.Generator(device="cuda"); g.manual_seed(100 + r)
    if name == "uniform":
        keys = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    else:
        u = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
        rank = torch.exp(u * float(np.log(1 << 24))).to(torch.int64).clamp_(1, 1 << 24); del u
        keys = rank * -7046029254386353131; del rank
    pay = (torch.arange(n, dtype=torch.int64, device="cuda") + r * n).to(torch.int32)
    for _ in range(2):
        cdist.sample_sort(keys, pay, ops, 64, gidx0=r * n)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    a.record()
    for _ in range(steps):
        k, p, info = cdist.sample_sort(keys, pay, ops, 64, gidx0=r * n)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    k, p, info = cdist.sample_sort(keys, pay, ops, 64, gidx0=r * n, profile=True)
    uk = k ^ (-2**63)
    ok = bool((uk[1:] >= uk[:-1]).all().item())
    eq = uk[1:] == uk[:-1]
    pl = p.to(torch.int64) & 0xFFFFFFFF
    stable = bool((pl[1:][eq] > pl[:-1][eq]).all().item()) if P * n <= (1 << 32) else None
    edges = torch.zeros(2 * P, dtype=torch.int64, device="cuda")
    edges[2 * r], edges[2 * r + 1] = uk[0], uk[-1]
    dist.all_reduce(edges)
    cnt = torch.tensor([k.numel()], dtype=torch.int64, device="cuda"); dist.all_reduce(cnt)
    e = edges.tolist()
    ok = ok and all(e[2 * i + 1] <= e[2 * i + 2] for i in range(P - 1)) and int(cnt.item()) == P * n
    okt = torch.tensor([1 if (ok and stable is not False) else 0], device="cuda"); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if r == 0:
        ms = float(t.item())
        print(json.dumps({"config": "C3 key-value sample sort, 2^%d (u64 key + u32 payload) per GPU" % int(np.log2(n)), "n_gpus": P,
                          "keys": name, "ms_per_step": round(ms, 3), "gpairs_per_s": round(P * n / ms / 1e6, 2),
                          "sorted_and_stable": bool(okt.item()), "fused_peer_scatter": bool(info.get("fused")),
                          "received_rank0": info["received"], "phases_ms_rank0": {x: round(y, 3) for x, y in info["phases_ms"].items()}}), flush=True)
    del keys, pay, k, p, uk, eq, pl
ops.close(); q.destroy(); ctx.destroy()
dist.barrier(); dist.destroy_process_group()
