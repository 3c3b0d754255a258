"""tests/dist_gpu_worker.py -- run under torchrun by test_sample_sort_two_gpus_fused_exchange:
sample sort over P GPUs (fused partition + CUDA-IPC peer exchange), keys-only u32 and
u64 keys + u32 payload, uniform and heavily duplicated, checked against a stable torch sort of
the gathered input."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo  # noqa: E402
from cl_ops_b200 import dist as cdist  # noqa: E402


def main():
    r, P, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = clo.Context(lr)
    q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
    ok = True
    for key_type, kdt, bits, with_payload in ((clo.UINT, torch.int32, 32, False), (clo.ULONG, torch.int64, 64, True)):
        for dup in (False, True):
            n = 300000 + 1000 * r
            g = torch.Generator(device="cuda"); g.manual_seed(1000 * r + bits + dup)
            hi = 50 if dup else (2**31 - 1 if bits == 32 else 2**62)
            keys = torch.randint(-hi if not dup else 0, hi, (n,), dtype=kdt, device="cuda", generator=g)
            n_all = [300000 + 1000 * i for i in range(P)]
            g0 = sum(n_all[:r])
            payload = (torch.arange(n, dtype=torch.int64, device="cuda") + g0).to(torch.int32) if with_payload else None
            for impl in ("c", "py"):          # the C-ABI orchestrator (product) and the module's own steps
                ops = cdist.GpuOps(clo, ctx, q, key_type)
                # receive buffers too small for the uniform u32 case: the scatter must be a no-op and
                # the NCCL all-to-all-v path must take over
                tiny = (bits == 32 and not dup)
                ops.setup_peer_exchange(1000 if tiny else int(1.5 * max(n_all)), kdt, with_payload, impl=impl)
                for call in range(2):            # second call reuses the receive buffers
                    out_k, out_p, info = cdist.sample_sort(keys, payload, ops, bits, gidx0=None if call else g0)
                    assert bool(info.get("fused")) == (not tiny), "wrong exchange path"
                    assert tiny or info.get("impl", "py") == impl, "wrong orchestrator"
                    # reference: gather everything, stable sort by unsigned key
                    gk = [torch.empty(m, dtype=kdt, device="cuda") for m in n_all]
                    dist.all_gather(gk, keys)
                    allk = torch.cat(gk)
                    uk = (allk.to(torch.int64) & 0xFFFFFFFF) if bits == 32 else (allk ^ (-2**63))
                    order = torch.argsort(uk, stable=True)
                    cnts = torch.zeros(P, dtype=torch.int64, device="cuda"); cnts[r] = out_k.numel()
                    dist.all_reduce(cnts)
                    lo = int(cnts[:r].sum().item())
                    sl = order[lo:lo + out_k.numel()]
                    ok &= bool(torch.equal(out_k, allk[sl]))
                    if with_payload:
                        ok &= bool(torch.equal(out_p.to(torch.int64), sl.to(torch.int64)))   # payload = global index: stability
                    ok &= int(cnts.sum().item()) == sum(n_all)
                ops.close()
            # ---- the same sort through the C-ABI orchestrator (clo_dist_*, csrc/dist.cu): splitters
            # picked by the library's rank kernel, collectives through the communicator callbacks
            cd = clo.CloDist(ctx)
            cd.sort_setup(key_type, int(1.5 * max(n_all)), with_payload)
            gk = [torch.empty(m, dtype=kdt, device="cuda") for m in n_all]
            dist.all_gather(gk, keys)
            allk = torch.cat(gk)
            uk = (allk.to(torch.int64) & 0xFFFFFFFF) if bits == 32 else (allk ^ (-2**63))
            order = torch.argsort(uk, stable=True)
            cap_out = int(1.5 * max(n_all))
            out_k = torch.empty(cap_out, dtype=kdt, device="cuda")
            out_p = torch.empty(cap_out, dtype=torch.int32, device="cuda") if with_payload else None
            bk, bo = clo.Buffer.wrap_tensor(ctx, keys), clo.Buffer.wrap_tensor(ctx, out_k)
            bp = clo.Buffer.wrap_tensor(ctx, payload) if with_payload else None
            bpo = clo.Buffer.wrap_tensor(ctx, out_p) if with_payload else None
            for call in range(2):
                n_out = cd.sort(q, bk, bp, n, bo, bpo, cap_out, gidx0=None if call else g0)
                cnts = torch.zeros(P, dtype=torch.int64, device="cuda"); cnts[r] = n_out
                dist.all_reduce(cnts)
                lo = int(cnts[:r].sum().item())
                sl = order[lo:lo + n_out]
                ok &= bool(torch.equal(out_k[:n_out], allk[sl]))
                if with_payload:
                    ok &= bool(torch.equal(out_p[:n_out].to(torch.int64), sl.to(torch.int64)))
                ok &= int(cnts.sum().item()) == sum(n_all)
                sent, recv = cd.counts()
                ok &= sum(recv) == n_out and sum(sent) == n
                # balance: the (key, index) splitters keep every slice near the mean, duplicates or not
                ok &= n_out < 1.25 * max(n_all)
            # a receive capacity that cannot hold a slice: error on every rank, nothing written
            cd.sort_setup(key_type, 1000, with_payload)
            try:
                cd.sort(q, bk, bp, n, bo, bpo, cap_out)
                ok = False
            except clo.CloError as e:
                ok &= "receive" in e.message
            for b in (bk, bo, bp, bpo):
                if b is not None:
                    b.destroy()
            cd.destroy()
    # ---- clo_dist_scan_with_device_data: u32 -> u64 bit-exact, f32 within tolerance
    cd = clo.CloDist(ctx)
    n = (1 << 22) + 4096 * r
    n_all = [(1 << 22) + 4096 * i for i in range(P)]
    for et, st, tdt, odt in ((clo.UINT, clo.ULONG, torch.int32, torch.int64), (clo.FLOAT, clo.FLOAT, torch.float32, torch.float32)):
        g = torch.Generator(device="cuda"); g.manual_seed(77 + r)
        x = torch.rand(n, device="cuda", generator=g) if tdt == torch.float32 else torch.randint(0, 1 << 20, (n,), dtype=tdt, device="cuda", generator=g)
        y = torch.empty(n, dtype=odt, device="cuda")
        sc = clo.CloScan("blelloch", ctx, et, st)
        bi, bo = clo.Buffer.wrap_tensor(ctx, x), clo.Buffer.wrap_tensor(ctx, y)
        cd.scan(sc, q, bi, bo, n)
        gx = [torch.empty(m, dtype=tdt, device="cuda") for m in n_all]
        dist.all_gather(gx, x)
        allx = torch.cat(gx)
        lo = sum(n_all[:r])
        if tdt == torch.float32:
            ref = (torch.cumsum(allx.to(torch.float64), 0) - allx.to(torch.float64))[lo:lo + n]
            ok &= bool(((y.to(torch.float64) - ref).abs() <= 1e-5 * ref.abs() + 1e-3).all())
        else:
            ref = (torch.cumsum(allx.to(torch.int64), 0) - allx.to(torch.int64))[lo:lo + n]
            ok &= bool(torch.equal(y, ref))
        bi.destroy(); bo.destroy(); sc.destroy()
    first, count = clo.CloDist.rng_partition(1000003, r, P)
    spans = [None] * P
    dist.all_gather_object(spans, (first, count))
    ok &= spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(P - 1)) and spans[-1][0] + spans[-1][1] == 1000003
    cd.destroy()
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if r == 0:
        print("DIST GPU OK" if int(t.item()) else "DIST GPU MISMATCH", flush=True)
    q.destroy(); ctx.destroy()
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
