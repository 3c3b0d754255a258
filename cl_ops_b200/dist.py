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
import os

import torch
import torch.distributed as dist

_SIGN64 = -0x8000000000000000
_I64_MAX = 0x7FFFFFFFFFFFFFFF


def _unsigned_order_i64(keys, key_bits):
    """int64 values whose signed order equals the unsigned order of the raw key bits."""
    if key_bits == 32:
        return keys.to(torch.int64) & 0xFFFFFFFF
    return keys.to(torch.int64) ^ _SIGN64


class _DevMem:
    """library-owned device memory seen as a torch tensor (zero copy, __cuda_array_interface__)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class GpuOps:
    """Local operators on CUDA tensors through the C-ABI (cl_ops_b200.lib())."""

    def __init__(self, clo, ctx, queue, key_type):
        self.clo, self.ctx, self.queue, self.key_type = clo, ctx, queue, key_type
        self.sorter = clo.CloSort("satradix", ctx, key_type)
        self.peer = None

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

    # ---- fused partition + exchange over peer memory (one process per GPU, one box)
    def setup_peer_exchange(self, capacity, key_dtype, with_payload, group=None, impl=None):
        """Allocate this rank's receive buffers (capacity elements), export them with CUDA IPC
        and map every peer's: afterwards the partition's scatter kernel writes each bucket
        straight into its destination rank over NVLink.  Collective; call once.

        impl "c" (default; CLO_DIST_IMPL overrides): the whole sort is ONE C-ABI call,
        clo_dist_sort_with_device_data (csrc/dist.cu) -- samples, splitters, sizes, scatter,
        barrier and local sort are orchestrated by the library, torch.distributed only serves
        its three collective callbacks.  impl "py": the same steps driven from this module
        (what the gloo tests exercise on the CPU)."""
        impl = impl or os.environ.get("CLO_DIST_IMPL", "c")
        if impl == "c":
            try:
                self.cd = self.clo.CloDist(self.ctx, group)
                self.cd.sort_setup(self.key_type, capacity, with_payload)
            except self.clo.CloError as e:
                # the library agrees on set-up failures across ranks (every rank gets the error), so
                # every rank takes the module's own driver below; say so, this is not the product path
                import sys
                print("cl_ops_b200.dist: clo_dist set-up failed (%s); using the Python driver" % e.message, file=sys.stderr)
                if getattr(self, "cd", None) is not None:
                    self.cd.destroy()
                self.cd = None
                impl = "py"
        if impl == "c":
            dev = torch.device("cuda", torch.cuda.current_device())
            self.cd_with_payload = with_payload
            self.cd_capacity = capacity
            self.cd_out_k = torch.empty(max(1, capacity), dtype=key_dtype, device=dev)
            self.cd_out_p = torch.empty(max(1, capacity), dtype=torch.int32, device=dev) if with_payload else None
            self.cd_bufs = (self._buf(self.cd_out_k), self._buf(self.cd_out_p) if with_payload else None)
            return
        P, r = dist.get_world_size(group), dist.get_rank(group)
        if P > 16:
            raise ValueError("the peer exchange supports at most 16 ranks per box (got %d)" % P)
        kb = torch.empty(0, dtype=key_dtype).element_size()
        own = [self.clo.Buffer(self.ctx, size=max(1, capacity) * kb)]
        if with_payload:
            own.append(self.clo.Buffer(self.ctx, size=max(1, capacity) * 4))
        handles = [None] * P
        dist.all_gather_object(handles, [b.ipc_export() for b in own], group=group)
        maps = []
        for i in range(P):
            maps.append(own if i == r else
                        [self.clo.Buffer.ipc_import(self.ctx, h, b.size) for h, b in zip(handles[i], own)])
        dev = torch.device("cuda", torch.cuda.current_device())
        ptrs = torch.zeros(2, 16, dtype=torch.int64)
        for i in range(P):
            for j, b in enumerate(maps[i]):
                ptrs[j, i] = b.ptr
        self.peer = {
            "P": P, "r": r, "capacity": capacity, "own": own, "maps": maps, "ptrs": ptrs.to(dev),
            "recv_k": torch.as_tensor(_DevMem(own[0].ptr, own[0].size), device=dev).view(key_dtype),
            "recv_p": torch.as_tensor(_DevMem(own[1].ptr, own[1].size), device=dev).view(torch.int32) if with_payload else None,
            "with_payload": with_payload, "group": group,
        }
        dist.barrier(group=group)

    def partition_count(self, keys, gidx0, splitter_keys, splitter_idx, nparts):
        counts = torch.zeros(nparts, dtype=torch.int64, device=keys.device)
        bk, bc = self._buf(keys), self._buf(counts)
        bsk = self._buf(splitter_keys) if nparts > 1 else None
        bsi = self._buf(splitter_idx) if nparts > 1 else None
        self.sorter.partition_count_with_device_data(self.queue, bk, keys.numel(), gidx0, bsk, bsi, nparts, bc)
        for b in (bk, bc, bsk, bsi):
            if b is not None:
                b.destroy()
        return counts

    def partition_scatter(self, keys, payload, gidx0, splitter_keys, splitter_idx, nparts, first_slot, ok):
        pe = self.peer
        bk = self._buf(keys)
        bp = self._buf(payload) if payload is not None else None
        bsk = self._buf(splitter_keys) if nparts > 1 else None
        bsi = self._buf(splitter_idx) if nparts > 1 else None
        bfs, bok = self._buf(first_slot), self._buf(ok)
        bd = self._buf(pe["ptrs"][0])
        bpd = self._buf(pe["ptrs"][1]) if payload is not None else None
        self.sorter.partition_scatter_with_device_data(self.queue, bk, bp, keys.numel(), gidx0, bsk, bsi, nparts,
                                                       bfs, bd, bpd, bok)
        for b in (bk, bp, bsk, bsi, bfs, bok, bd, bpd):
            if b is not None:
                b.destroy()

    def sort_received(self, n_recv):
        """stable sort of the first n_recv received elements into fresh tensors (the receive
        buffers themselves are overwritten by the peers in the next exchange)"""
        pe = self.peer
        if pe["with_payload"]:
            k, p = self.sort(pe["recv_k"][:n_recv], pe["recv_p"][:n_recv])
            return k.clone(), p.clone()
        out = torch.empty(n_recv, dtype=pe["recv_k"].dtype, device=pe["recv_k"].device)
        if n_recv:
            bi, bo = self._buf(pe["recv_k"][:n_recv]), self._buf(out)
            self.sorter.with_device_data(self.queue, bi, bo, n_recv)
            bi.destroy(); bo.destroy()
        return out, None

    def sort_c(self, keys, payload, gidx0, profile):
        """the C-ABI sample sort; returns (keys, payload, info) like sample_sort, or None when a
        receive buffer is too small (the library then moved nothing)"""
        cd = self.cd
        cd.set_timing(bool(profile))
        bk = self._buf(keys)
        bp = self._buf(payload) if payload is not None else None
        try:
            n_out = cd.sort(self.queue, bk, bp, keys.numel(), self.cd_bufs[0], self.cd_bufs[1], self.cd_capacity, gidx0)
        except self.clo.CloError as e:
            if "receive" in e.message:
                return None
            raise
        finally:
            bk.destroy()
            if bp is not None:
                bp.destroy()
        sent, recv = cd.counts()
        r = dist.get_rank(cd.group)
        info = {"sent": keys.numel() - sent[r], "received": n_out, "gidx0": gidx0, "fused": True, "impl": "c",
                "send_counts": sent, "recv_counts": recv, "phases_ms": cd.phases_ms() if profile else {}}
        out_p = self.cd_out_p[:n_out] if payload is not None else None
        return self.cd_out_k[:n_out], out_p, info

    def close(self):
        if getattr(self, "cd", None) is not None:
            for b in self.cd_bufs:
                if b is not None:
                    b.destroy()
            self.cd.destroy()
            self.cd = None
            self.cd_out_k = self.cd_out_p = None
        if self.peer:
            dist.barrier(group=self.peer["group"])
            for i, bl in enumerate(self.peer["maps"]):
                if i != self.peer["r"]:
                    for b in bl:
                        b.destroy()
            self.peer["recv_k"] = self.peer["recv_p"] = None
            dist.barrier(group=self.peer["group"])
            for b in self.peer["own"]:
                b.destroy()
            self.peer = None
        self.sorter.destroy()


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


def _exchange(src, dst, send_list, recv_list, r, P, group):
    """all-to-all-v of contiguous buckets.  The rank's own bucket never touches NCCL (a
    device-local copy); the others go as one group of NCCL sends and receives over NVLink
    (CLO_DIST_EXCHANGE=a2a selects the single all_to_all_single call instead)."""
    import os
    if os.environ.get("CLO_DIST_EXCHANGE") == "a2a" or not src.is_cuda:
        dist.all_to_all_single(dst, src, output_split_sizes=recv_list, input_split_sizes=send_list, group=group)
        return
    so = [0] * (P + 1)
    ro = [0] * (P + 1)
    for i in range(P):
        so[i + 1] = so[i] + send_list[i]
        ro[i + 1] = ro[i] + recv_list[i]
    ops = []
    for k in range(1, P):
        to, fr = (r + k) % P, (r - k) % P
        if send_list[to]:
            ops.append(dist.P2POp(dist.isend, src[so[to]:so[to + 1]], dist.get_global_rank(group, to) if group else to, group=group))
        if recv_list[fr]:
            ops.append(dist.P2POp(dist.irecv, dst[ro[fr]:ro[fr + 1]], dist.get_global_rank(group, fr) if group else fr, group=group))
    works = dist.batch_isend_irecv(ops) if ops else []
    dst[ro[r]:ro[r + 1]].copy_(src[so[r]:so[r + 1]])
    for w in works:
        w.wait()


class _Graphed:
    """A pure device function of STATIC input tensors, captured once as a CUDA graph and
    replayed: the dozen small kernels of the splitter selection (and of the bucket-size
    bookkeeping) cost one launch instead of twelve.  Plumbing only."""

    def __init__(self, fn, inputs):
        self.inputs = inputs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                fn(*inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.outputs = fn(*inputs)

    def __call__(self):
        self.graph.replay()
        return self.outputs


def _pick_splitters(allr, lane_idx, pick_mul, P, cap, key_bits):
    """P-1 splitters (key, global index) from the gathered sample rows; device only."""
    n_all = allr[:, 0]
    g0_all = torch.cumsum(n_all, 0) - n_all
    valid = lane_idx < allr[:, 1:2]
    all_k = torch.where(valid, allr[:, 2:2 + cap], _I64_MAX).reshape(-1)
    all_i = (allr[:, 2 + cap:] + g0_all[:, None]).reshape(-1)
    order = torch.argsort(all_k, stable=True)
    total = allr[:, 1].sum()
    pick = (pick_mul * total) // P
    sel = order[pick]
    spl_k64, spl_i = all_k[sel], all_i[sel]
    if key_bits == 32:
        spl_keys = spl_k64.to(torch.int32)            # low 32 bits = the raw u32 pattern
    else:
        spl_keys = spl_k64 ^ _SIGN64
    return spl_keys.contiguous(), spl_i.contiguous(), g0_all


def _slots_from_sizes(M, r, P, capacity):
    """M[src][dst] bucket sizes -> (first slot of my bucket in every destination, fits flag,
    [what I receive from each source, what I send to each destination, fits])"""
    first_slot = torch.zeros(16, dtype=torch.int64, device=M.device)
    first_slot[:P] = (torch.cumsum(M, 0) - M)[r]
    ok = (M.sum(0) <= capacity).all().to(torch.int32).reshape(1)
    host = torch.cat([M[:, r], M[r], ok.to(torch.int64)])
    return first_slot, ok, host


def sample_sort(keys, payload, ops, key_bits, group=None, samples_per_rank=None, profile=False, gidx0=None):
    """Globally stable sort of the concatenation of every rank's `keys` (rank order).

    keys: 1-D int32 (u32 bit pattern) or int64 (u64 bit pattern) tensor on this rank.
    payload: optional int32 tensor, moved with the keys.
    gidx0: global index of this rank's first element when the caller knows it (saves one
    host synchronisation); None = derive it from the gathered element counts.
    Returns (keys, payload, info): rank r ends up with the r-th slice of the sorted
    sequence (slice lengths differ slightly between ranks).

    Host synchronisations per call: one (the bucket sizes NCCL needs as host integers), plus
    one more when gidx0 is not given.
    """
    P = dist.get_world_size(group)
    r = dist.get_rank(group)
    dev = keys.device
    n_local = keys.numel()
    ph = _Phases(profile and keys.is_cuda)
    ph.mark("start")

    if P == 1:
        k, p = ops.sort(keys, payload)
        return k, p, {"sent": 0, "received": n_local, "gidx0": 0}

    # the product path on GPUs: one call into the library (clo_dist_sort_with_device_data)
    if getattr(ops, "cd", None) is not None and ops.cd_with_payload == (payload is not None):
        res = ops.sort_c(keys, payload, gidx0, profile)
        if res is not None:
            return res
        # a receive buffer would overflow (nothing was written): the NCCL all-to-all-v path below

    # 1) regular samples of the unsorted local data.  One all-gather carries, per rank:
    #    [n_local, s, sample keys (cap), sample positions (cap)]; global indices are formed
    #    after the gather from the gathered counts, so nothing has to come back to the host.
    cap = samples_per_rank or 64 * P
    s = min(cap, n_local)
    # the header and the sample positions only depend on (n_local, cap): built once
    cache = getattr(ops, "_sample_cache", None)
    if cache is None:
        cache = {}
        try:
            ops._sample_cache = cache
        except Exception:
            pass
    ck = (n_local, cap, str(dev))
    if ck not in cache:
        pos = (torch.arange(s, dtype=torch.int64, device=dev) * n_local) // max(s, 1)
        pad = torch.zeros(cap - s, dtype=torch.int64, device=dev)
        cache[ck] = (torch.tensor([n_local, s], dtype=torch.int64, device=dev), pos, pad,
                     torch.arange(cap, dtype=torch.int64, device=dev)[None, :],
                     torch.arange(1, P, dtype=torch.int64, device=dev))
    hdr, pos, pad, lane_idx, pick_mul = cache[ck]
    row = torch.cat([hdr, _unsigned_order_i64(keys[pos], key_bits), pad, pos, pad])
    use_graphs = keys.is_cuda and os.environ.get("CLO_DIST_GRAPHS", "1") != "0"
    gk = ("spl", P, cap, key_bits, str(dev))
    if use_graphs and gk in cache:
        allr, spl_graph = cache[gk]
    else:
        allr = torch.empty(P, 2 + 2 * cap, dtype=torch.int64, device=dev)
        spl_graph = None
    dist.all_gather_into_tensor(allr.view(-1), row, group=group)
    ph.mark("samples+allgather")

    # 2) splitters (identical on every rank: same data, same deterministic procedure).
    #    The gathered samples are already in global-index order (rank, then position), so ONE
    #    stable sort by key gives the lexicographic (key, index) order; unused slots carry the
    #    largest key and sort behind everything.  All on the device: no host round trip.
    if use_graphs:
        if spl_graph is None:
            spl_graph = _Graphed(lambda a: _pick_splitters(a, lane_idx, pick_mul, P, cap, key_bits), [allr])
            cache[gk] = (allr, spl_graph)
        spl_keys, spl_i, g0_all = spl_graph()
    else:
        spl_keys, spl_i, g0_all = _pick_splitters(allr, lane_idx, pick_mul, P, cap, key_bits)
    if gidx0 is None:
        gidx0 = int(g0_all[r].item())
    ph.mark("splitters")

    # 3+4 fused) count -> all-gather of the P x P bucket sizes -> every bucket is scattered
    #    straight into the receive buffer of its destination rank (peer memory, NVLink)
    pe = getattr(ops, "peer", None)
    if pe is not None and pe["with_payload"] == (payload is not None):
        counts = ops.partition_count(keys, gidx0, spl_keys, spl_i, P)
        ph.mark("count")
        mk = ("sizes", P, r, pe["capacity"], str(dev))
        if use_graphs and mk in cache:
            M, size_graph = cache[mk]
        else:
            M, size_graph = torch.empty(P, P, dtype=torch.int64, device=dev), None
        dist.all_gather_into_tensor(M.view(-1), counts, group=group)        # M[src][dst]
        if use_graphs:
            if size_graph is None:
                size_graph = _Graphed(lambda m: _slots_from_sizes(m, r, P, pe["capacity"]), [M])
                cache[mk] = (M, size_graph)
            first_slot, ok, host_vec = size_graph()
        else:
            first_slot, ok, host_vec = _slots_from_sizes(M, r, P, pe["capacity"])
        ph.mark("sizes all-gather")
        ops.partition_scatter(keys, payload, gidx0, spl_keys, spl_i, P, first_slot, ok)
        ph.mark("scatter to peers")
        done = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(done, group=group)                 # every peer's writes have landed
        host = host_vec.tolist()                           # the one host synchronisation
        recv_list, send_list, fits = host[:P], host[P:2 * P], bool(host[2 * P])
        ph.mark("barrier+sizes to host")
        if fits:
            n_recv = sum(recv_list)
            out_k, out_p = ops.sort_received(n_recv)
            ph.mark("local sort")
            info = {"sent": n_local - send_list[r], "received": n_recv, "gidx0": gidx0, "fused": True,
                    "send_counts": send_list, "recv_counts": recv_list, "phases_ms": ph.result()}
            return out_k, out_p, info
        # a receive buffer would overflow (nothing was written): fall through to the NCCL path

    # 3) stable local partition into P buckets; bucket sizes
    part_k, part_p, counts = ops.partition(keys, payload, gidx0, spl_keys, spl_i, P)
    ph.mark("partition")

    # 4) exchange: counts, then the buckets (all-to-all-v)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    both = torch.stack([counts, recv_counts]).tolist()          # the one host synchronisation
    send_list = [int(x) for x in both[0]]
    recv_list = [int(x) for x in both[1]]
    n_recv = sum(recv_list)
    ph.mark("counts")
    recv_k = torch.empty(n_recv, dtype=keys.dtype, device=dev)
    recv_p = torch.empty(n_recv, dtype=payload.dtype, device=dev) if payload is not None else None
    ph.mark("alloc")
    _exchange(part_k, recv_k, send_list, recv_list, r, P, group)
    if payload is not None:
        _exchange(part_p, recv_p, send_list, recv_list, r, P, group)

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
