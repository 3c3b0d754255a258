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
            ops = cdist.GpuOps(clo, ctx, q, key_type)
            # receive buffers too small for the uniform u32 case: the scatter must be a no-op and
            # the NCCL all-to-all-v path must take over
            tiny = (bits == 32 and not dup)
            ops.setup_peer_exchange(1000 if tiny else int(1.5 * max(n_all)), kdt, with_payload)
            for call in range(2):            # second call reuses the receive buffers
                out_k, out_p, info = cdist.sample_sort(keys, payload, ops, bits, gidx0=None if call else g0)
                assert bool(info.get("fused")) == (not tiny), "wrong exchange path"
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
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if r == 0:
        print("DIST GPU OK" if int(t.item()) else "DIST GPU MISMATCH", flush=True)
    q.destroy(); ctx.destroy()
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
