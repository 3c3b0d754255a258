"""cl_ops_b200.dist -- the three operators sharded over the GPUs of one box.

One process per GPU (torchrun); `torch.distributed` (NCCL over NVLink/NVSwitch on the
GPUs, gloo in the CPU tests) is only the plumbing.  The reference is single-device
(/root/reference/src/cl_ops/sort/clo_sort_abstract.c:335 "first device in context"), so
everything here is new; the partitioning follows SURVEY.md section 8(e):

* sort  -- sample sort: regular samples (key, global index) -> all-gather -> P-1 splitters
           -> stable local partition by splitter (clo_sort_partition_with_device_data)
           -> all-to-all-v of keys (+ payload) -> stable local LSD radix sort.
           Splitters are tie-broken on the global index, so heavy duplication (Zipf) stays
           balanced and the result is globally stable.
* scan  -- per-GPU total (clo_scan_reduce_with_device_data) -> all-gather of P totals ->
           per-GPU exclusive scan with a device-resident carry-in
           (clo_scan_with_device_data_carry).  Integer results are bit-exact.
* rng   -- stream partitioning, no communication: rank r owns work-items
           [r*G/P, (r+1)*G/P) (clo_rng_new_dev_gid_offset).

The local operators are injected (`ops`), so the host-side logic can be exercised with
gloo on CPU tensors by the tests; the product binding is `GpuOps`, which calls the C-ABI
library and has no fallback.
"""
import torch
import torch.distributed as dist

_SIGN64 = -0x8000000000000000


def _unsigned_order_i64(keys, key_bits):
    """int64 values whose signed order equals the unsigned order of the raw key bits."""
    if key_bits == 32:
        return keys.to(torch.int64) & 0xFFFFFFFF
    return keys.to(torch.int64) ^ _SIGN64


class GpuOps:
    """Local operators on CUDA tensors through the C-ABI (cl_ops_b200.lib())."""

    def __init__(self, clo, ctx, queue, key_type):
        self.clo, self.ctx, self.queue, self.key_type = clo, ctx, queue, key_type
        self.sorter = clo.CloSort("satradix", ctx, key_type)

    def _buf(self, t):
        return self.clo.Buffer.wrap_tensor(self.ctx, t)

    def partition(self, keys, payload, gidx0, splitter_keys, splitter_idx, nparts):
        n = keys.numel()
        keys_out = torch.empty_like(keys)
        payload_out = torch.empty_like(payload) if payload is not None else None
        counts = torch.zeros(nparts, dtype=torch.int64, device=keys.device)
        bufs = [self._buf(keys), self._buf(keys_out), self._buf(counts)]
        bp = self._buf(payload) if payload is not None else None
        bpo = self._buf(payload_out) if payload is not None else None
        bsk = self._buf(splitter_keys) if nparts > 1 else None
        bsi = self._buf(splitter_idx) if nparts > 1 else None
        self.sorter.partition_with_device_data(self.queue, bufs[0], bp, bufs[1], bpo, n, gidx0,
                                               bsk, bsi, nparts, bufs[2])
        for b in bufs + [x for x in (bp, bpo, bsk, bsi) if x is not None]:
            b.destroy()
        return keys_out, payload_out, counts

    def sort(self, keys, payload):
        n = keys.numel()
        if n == 0:
            return keys, payload
        bk = self._buf(keys)
        if payload is None:
            self.sorter.with_device_data(self.queue, bk, None, n)
        else:
            bp = self._buf(payload)
            self.sorter.pairs_with_device_data(self.queue, bk, bp, n)
            bp.destroy()
        bk.destroy()
        return keys, payload

    def close(self):
        self.sorter.destroy()


def choose_splitters(sample_keys_i64, sample_idx, nparts):
    """P-1 splitters from the gathered samples: lexicographic (key, global index) order,
    regular positions.  `sample_keys_i64` is in unsigned-order int64 form."""
    total = sample_keys_i64.numel()
    # lexicographic sort: stable sort by index, then stable sort by key
    o1 = torch.argsort(sample_idx, stable=True)
    k1 = sample_keys_i64[o1]
    o2 = torch.argsort(k1, stable=True)
    order = o1[o2]
    pos = torch.tensor([(k * total) // nparts for k in range(1, nparts)], dtype=torch.int64,
                       device=sample_keys_i64.device)
    sel = order[pos]
    return sel


class _Phases:
    """optional per-phase device timing (CUDA events) for profiling the sample sort"""

    def __init__(self, enabled):
        self.enabled, self.marks = enabled, []

    def mark(self, name):
        if self.enabled:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def result(self):
        if not self.enabled or len(self.marks) < 2:
            return {}
        torch.cuda.synchronize()
        return {self.marks[i + 1][0]: self.marks[i][1].elapsed_time(self.marks[i + 1][1])
                for i in range(len(self.marks) - 1)}


def sample_sort(keys, payload, ops, key_bits, group=None, samples_per_rank=None, profile=False):
    """Globally stable sort of the concatenation of every rank's `keys` (rank order).

    keys: 1-D int32 (u32 bit pattern) or int64 (u64 bit pattern) tensor on this rank.
    payload: optional int32 tensor, moved with the keys.
    Returns (keys, payload, info): rank r ends up with the r-th slice of the sorted
    sequence (slice lengths differ slightly between ranks).
    """
    P = dist.get_world_size(group)
    r = dist.get_rank(group)
    dev = keys.device
    n_local = keys.numel()
    ph = _Phases(profile and keys.is_cuda)
    ph.mark("start")

    # global index of my first element (ranks may hold different counts)
    n_all = torch.zeros(P, dtype=torch.int64, device=dev)
    n_all[r] = n_local
    dist.all_reduce(n_all, group=group)
    gidx0 = int(n_all[:r].sum().item())

    if P == 1:
        k, p = ops.sort(keys, payload)
        return k, p, {"sent": 0, "received": n_local, "gidx0": 0}

    # 1) regular samples of the unsorted local data, with their global indices
    s = samples_per_rank or 64 * P
    s = max(1, min(s, n_local)) if n_local > 0 else 0
    smax = torch.tensor([s], dtype=torch.int64, device=dev)
    dist.all_reduce(smax, op=dist.ReduceOp.MAX, group=group)
    smax = int(smax.item())
    samp_k = torch.zeros(smax, dtype=torch.int64, device=dev)
    samp_i = torch.full((smax,), -1, dtype=torch.int64, device=dev)   # -1 marks padding
    if s > 0:
        pos = (torch.arange(s, dtype=torch.int64, device=dev) * n_local) // s
        samp_k[:s] = _unsigned_order_i64(keys[pos], key_bits)
        samp_i[:s] = gidx0 + pos
    all_k = torch.empty(P * smax, dtype=torch.int64, device=dev)
    all_i = torch.empty(P * smax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_k, samp_k, group=group)
    dist.all_gather_into_tensor(all_i, samp_i, group=group)
    valid = all_i >= 0
    all_k, all_i = all_k[valid], all_i[valid]

    ph.mark("samples+allgather")
    # 2) splitters (identical on every rank: same data, same deterministic procedure)
    sel = choose_splitters(all_k, all_i, P)
    spl_k64, spl_i = all_k[sel], all_i[sel]
    if key_bits == 32:
        spl_keys = spl_k64.to(torch.int32)            # low 32 bits = the raw u32 pattern
    else:
        spl_keys = spl_k64 ^ _SIGN64
    spl_keys, spl_i = spl_keys.contiguous(), spl_i.contiguous()

    ph.mark("splitters")
    # 3) stable local partition into P buckets; bucket sizes
    part_k, part_p, counts = ops.partition(keys, payload, gidx0, spl_keys, spl_i, P)
    ph.mark("partition")

    # 4) exchange: counts, then the buckets (all-to-all-v)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    send_list = [int(x) for x in counts.tolist()]
    recv_list = [int(x) for x in recv_counts.tolist()]
    n_recv = sum(recv_list)
    recv_k = torch.empty(n_recv, dtype=keys.dtype, device=dev)
    dist.all_to_all_single(recv_k, part_k, output_split_sizes=recv_list, input_split_sizes=send_list,
                           group=group)
    recv_p = None
    if payload is not None:
        recv_p = torch.empty(n_recv, dtype=payload.dtype, device=dev)
        dist.all_to_all_single(recv_p, part_p, output_split_sizes=recv_list, input_split_sizes=send_list,
                               group=group)

    ph.mark("exchange")
    # 5) stable local sort of what arrived (chunks are in source-rank order, each in its
    #    original order, so a stable sort keeps the global order of equal keys)
    out_k, out_p = ops.sort(recv_k, recv_p)
    ph.mark("local sort")
    info = {"sent": n_local - send_list[r], "received": n_recv, "gidx0": gidx0,
            "send_counts": send_list, "recv_counts": recv_list, "phases_ms": ph.result()}
    return out_k, out_p, info


def exclusive_offsets(totals):
    """Carry-in of every rank from the all-gathered per-rank totals (wrap-around kept)."""
    return torch.cumsum(totals, 0) - totals


def dist_scan(local_reduce, local_scan_with_carry, data, sum_dtype, group=None):
    """Exclusive scan of the concatenation of every rank's `data`.

    local_reduce(data) -> 1-element tensor (sum type) with this rank's total;
    local_scan_with_carry(data, carry_1elem_tensor) -> scanned tensor.
    Exchange: one all-gather of P totals.
    """
    P = dist.get_world_size(group)
    r = dist.get_rank(group)
    total = local_reduce(data)
    if P == 1:
        return local_scan_with_carry(data, torch.zeros_like(total))
    totals = torch.empty(P, dtype=total.dtype, device=total.device)
    dist.all_gather_into_tensor(totals, total, group=group)
    if totals.dtype.is_floating_point:
        carry = exclusive_offsets(totals.to(torch.float64))[r:r + 1].to(sum_dtype)
    else:
        carry = exclusive_offsets(totals)[r:r + 1].to(sum_dtype)
    return local_scan_with_carry(data, carry.contiguous())


def rng_partition(total_streams, group=None):
    """(first stream, stream count) of this rank: contiguous, no communication."""
    P = dist.get_world_size(group) if dist.is_initialized() else 1
    r = dist.get_rank(group) if dist.is_initialized() else 0
    base, rem = divmod(total_streams, P)
    first = r * base + min(r, rem)
    return first, base + (1 if r < rem else 0)
