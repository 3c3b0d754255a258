import sys, os, torch, numpy as np
sys.path.insert(0, "/root/repo")
import cl_ops_b200 as clo
ctx = clo.Context(); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
n = 1 << 20
a0 = torch.randint(-2**31, 2**31 - 1, (n,), dtype=torch.int32, device="cuda")
t = torch.empty_like(a0); b = clo.Buffer.wrap_tensor(ctx, t)
def timed(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    x, y = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x.record()
    for _ in range(it): fn()
    y.record(); torch.cuda.synchronize()
    return x.elapsed_time(y) / it
tc = timed(lambda: t.copy_(a0))
for alg, opts in (("sbitonic", None), ("abitonic", "maxps=4"), ("abitonic", "maxps=3"), ("abitonic", "maxps=4,maxsfs=9")):
    s = clo.CloSort(alg, ctx, clo.UINT, options=opts)
    def run():
        t.copy_(a0); s.with_device_data(q, b, None, n)
    ms = timed(run) - tc
    u = t.to(torch.int64) & 0xFFFFFFFF
    print(alg, opts, "ms %.4f" % ms, "sorted", bool((u[1:] >= u[:-1]).all().item()), flush=True)
    s.destroy()
