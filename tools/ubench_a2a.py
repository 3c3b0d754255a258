"""tools/ubench_a2a.py -- how fast does this box move the sample sort's exchange?
torchrun --nproc-per-node N tools/ubench_a2a.py : all_to_all_single vs remote-only send/recv."""
import os, torch, torch.distributed as dist
r = int(os.environ["RANK"]); P = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 28
src = torch.arange(n, dtype=torch.int32, device="cuda")
dst = torch.empty_like(src)
per = n // P
splits = [per] * P

def timed(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it

t1 = timed(lambda: dist.all_to_all_single(dst, src, output_split_sizes=splits, input_split_sizes=splits))
def remote_only():
    ops = []
    for k in range(1, P):
        to, fr = (r + k) % P, (r - k) % P
        ops.append(dist.P2POp(dist.isend, src[to * per:(to + 1) * per], to))
        ops.append(dist.P2POp(dist.irecv, dst[fr * per:(fr + 1) * per], fr))
    for w in dist.batch_isend_irecv(ops): w.wait()
    dst[r * per:(r + 1) * per].copy_(src[r * per:(r + 1) * per])
t2 = timed(remote_only)
t3 = timed(lambda: dst.copy_(src))
def fresh_alloc():
    global dst
    dst = torch.empty(n + 1024 * (r + 1), dtype=torch.int32, device="cuda")
    remote_only()
t4 = timed(fresh_alloc)
def with_sync():
    torch.cuda.synchronize()
    remote_only()
t5 = timed(with_sync)
uneven = [per + 4096 * (1 if (i + r) % 2 else -1) for i in range(P)]
uneven[-1] = n - sum(uneven[:-1])
# what rank i sends to me = its uneven[(me)] ; compute via all_to_all of counts
cnt = torch.tensor(uneven, dtype=torch.int64, device="cuda"); rc = torch.empty_like(cnt)
dist.all_to_all_single(rc, cnt); rl = rc.tolist()
dst2 = torch.empty(sum(rl), dtype=torch.int32, device="cuda")
t6 = timed(lambda: dist.all_to_all_single(dst2, src, output_split_sizes=rl, input_split_sizes=uneven))
remote_bytes = 4 * per * (P - 1)
if r == 0:
    print("P=%d all_to_all_single %.3f ms (%.0f GB/s remote per GPU per direction) | send/recv remote only + local copy %.3f ms (%.0f GB/s) | local 1 GiB copy %.3f ms | p2p with fresh dst alloc %.3f ms | p2p after host sync %.3f ms | a2a uneven splits %.3f ms" %
          (P, t1, remote_bytes / t1 / 1e6, t2, remote_bytes / t2 / 1e6, t3, t4, t5, t6), flush=True)
dist.barrier(); dist.destroy_process_group()
