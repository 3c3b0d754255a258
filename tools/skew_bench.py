"""tools/skew_bench.py -- sort throughput on skewed key distributions (2^28 u32 keys)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo

n = 1 << int(os.environ.get("LOG2N", "28"))
ctx = clo.Context(); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda"); g.manual_seed(1)
u = torch.rand(n, device="cuda", generator=g, dtype=torch.float32)
cases = {
    "uniform": torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda", generator=g),
    "zipf(1.0) over 2^20 values, hashed": None,
    "u^8 (dense small keys)": (u.double().pow(8) * 4294967295.0).to(torch.int64).to(torch.int32),
    "16 distinct values": (torch.randint(0, 16, (n,), device="cuda", generator=g, dtype=torch.int32) * 0x01010101),
    "all equal": torch.full((n,), 0x12345678, dtype=torch.int32, device="cuda"),
    "sorted": torch.arange(n, dtype=torch.int32, device="cuda"),
}
# zipf via inverse CDF approximation: rank = floor(exp(u * ln(V))) with V = 2^20 (s = 1)
V = 1 << 20
rank = torch.exp(u.double() * torch.log(torch.tensor(float(V), device="cuda", dtype=torch.float64))).to(torch.int64).clamp_(1, V)
cases["zipf(1.0) over 2^20 values, hashed"] = ((rank * 2654435761) & 0xFFFFFFFF).to(torch.int32)
del u, rank
s = clo.CloSort("satradix", ctx, clo.UINT)
out = torch.empty(n, dtype=torch.int32, device="cuda")
bo = clo.Buffer.wrap_tensor(ctx, out)
only = os.environ.get("CASES")
for name, t in cases.items():
    if only and not any(name.startswith(c) for c in only.split(",")):
        continue
    bi = clo.Buffer.wrap_tensor(ctx, t)
    for _ in range(2):
        s.with_device_data(q, bi, bo, n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        s.with_device_data(q, bi, bo, n)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    uo = out.to(torch.int64) & 0xFFFFFFFF
    ok = bool((uo[1:] >= uo[:-1]).all().item())
    d = s.debug(q)
    s.set_timing(True)
    s.with_device_data(q, bi, bo, n)
    torch.cuda.synchronize()
    tm = [round(x, 3) for x in s.get_timing()]
    s.set_timing(False)
    rec = {"keys": name, "ms": round(ms, 3), "gkeys": round(n / ms / 1e6, 1), "sorted": ok, "repaired": d[1], "timeout": d[0],
           "kernel_ms": tm}
    if os.environ.get("CLO_RADIX_PROFILE"):
        # cycles per phase of worker thread 0, summed over the tiles of the LAST call (radix_v6.cuh mark())
        names = ["P1 count+loadwait", "B1 wait", "P2 digits", "B2 wait", "P5 write-out", "B3 wait", "P3 place+P4 load"]
        tiles = 4 * ((n + 8191) // 8192)
        rec["cycles_per_tile"] = {k: round(v / tiles, 1) for k, v in zip(names, d[2:9])}
    print(json.dumps(rec), flush=True)
    bi.destroy(); del uo
