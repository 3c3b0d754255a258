"""CPU (gloo, world_size 2 and 3): host-side logic of the multi-GPU paths in
cl_ops_b200/dist.py -- splitter choice, count/bucket exchange, carry-in of the scan, stream
partitioning.  The local operators are test doubles written against the same contract as
the C-ABI entry points (clo_sort_partition_with_device_data, clo_sort_*_with_device_data);
the GPU tests check the real kernels against the oracle separately."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from cl_ops_b200 import dist as cdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class CpuOps:
    """Test double of GpuOps on CPU tensors (numpy)."""

    def __init__(self, key_bits):
        self.key_bits = key_bits
        self.udt = np.uint32 if key_bits == 32 else np.uint64

    def partition(self, keys, payload, gidx0, splitter_keys, splitter_idx, nparts):
        k = keys.numpy().view(self.udt)
        g = gidx0 + np.arange(k.size, dtype=np.uint64)
        bucket = np.zeros(k.size, dtype=np.int64)
        sk = splitter_keys.numpy().view(self.udt)
        si = splitter_idx.numpy().astype(np.uint64)
        for a, b in zip(sk, si):
            bucket += ((a < k) | ((a == k) & (b <= g))).astype(np.int64)
        order = np.argsort(bucket, kind="stable")
        counts = torch.from_numpy(np.bincount(bucket, minlength=nparts).astype(np.int64))
        pk = torch.from_numpy(k[order].view(keys.numpy().dtype).copy())
        pp = torch.from_numpy(payload.numpy()[order].copy()) if payload is not None else None
        return pk, pp, counts

    def sort(self, keys, payload):
        k = keys.numpy().view(self.udt)
        order = np.argsort(k, kind="stable")
        sk = torch.from_numpy(k[order].view(keys.numpy().dtype).copy())
        sp = torch.from_numpy(payload.numpy()[order].copy()) if payload is not None else None
        return sk, sp


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        if case == "sort32_dups":
            n = 5000 + 37 * rank
            keys = rng.integers(0, 40, size=n).astype(np.uint32)      # heavy duplication
            keys[::5] = rng.integers(0, 2**32, size=keys[::5].size, dtype=np.uint64).astype(np.uint32)
            payload = (rank * 1_000_000 + np.arange(n)).astype(np.int32)
            k, p, info = cdist.sample_sort(torch.from_numpy(keys.view(np.int32)), torch.from_numpy(payload),
                                           CpuOps(32), 32)
            q.put((rank, "in", keys, payload))
            q.put((rank, "out", k.numpy().view(np.uint32).copy(), p.numpy().copy()))
        elif case == "sort64":
            n = 3000
            keys = rng.integers(0, 2**64, size=n, dtype=np.uint64)
            k, p, info = cdist.sample_sort(torch.from_numpy(keys.view(np.int64)), None, CpuOps(64), 64)
            q.put((rank, "in", keys, None))
            q.put((rank, "out", k.numpy().view(np.uint64).copy(), None))
        elif case == "scan":
            n = 4096 + rank
            a = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)

            def red(x):
                return torch.tensor([int(x.numpy().view(np.uint32).astype(np.uint64).sum() & 0xFFFFFFFFFFFFFFFF)],
                                    dtype=torch.int64) if False else \
                    torch.from_numpy(np.array([x.numpy().view(np.uint32).astype(np.uint64).sum()],
                                              dtype=np.uint64).view(np.int64))

            def scan(x, carry):
                s = oracle.scan(x.numpy().view(np.uint32), oracle.UINT, oracle.ULONG)
                return torch.from_numpy((s + carry.numpy().view(np.uint64)[0]).view(np.int64))

            out = cdist.dist_scan(red, scan, torch.from_numpy(a.view(np.int32)), torch.int64)
            q.put((rank, "in", a, None))
            q.put((rank, "out", out.numpy().view(np.uint64).copy(), None))
        elif case == "rng":
            first, count = cdist.rng_partition(1000)
            q.put((rank, "out", np.array([first, count]), None))
    finally:
        dist.destroy_process_group()


def _run(case, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    msgs = [q.get(timeout=120) for _ in range(world * (1 if case == "rng" else 2))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return msgs


@pytest.mark.parametrize("world", [2, 3])
def test_sample_sort_is_globally_stable(world):
    msgs = _run("sort32_dups", world)
    ins = {r: (k, p) for r, tag, k, p in msgs if tag == "in"}
    outs = {r: (k, p) for r, tag, k, p in msgs if tag == "out"}
    keys = np.concatenate([ins[r][0] for r in range(world)])
    pay = np.concatenate([ins[r][1] for r in range(world)]).view(np.uint32)
    wk, wp = oracle.sort_pairs(keys, pay, oracle.UINT)
    gk = np.concatenate([outs[r][0] for r in range(world)])
    gp = np.concatenate([outs[r][1] for r in range(world)]).view(np.uint32)
    assert np.array_equal(gk, wk)
    assert np.array_equal(gp, wp), "equal keys lost their global input order"
    sizes = [outs[r][0].size for r in range(world)]
    assert max(sizes) < 1.5 * keys.size / world, "buckets unbalanced under duplication: %s" % sizes


def test_sample_sort_u64_keys_only():
    msgs = _run("sort64", 2)
    keys = np.concatenate([k for r, tag, k, p in sorted(msgs, key=lambda m: m[0]) if tag == "in"])
    out = np.concatenate([k for r, tag, k, p in sorted(msgs, key=lambda m: m[0]) if tag == "out"])
    assert np.array_equal(out, np.sort(keys))


def test_dist_scan_carry_in():
    msgs = _run("scan", 2)
    a = np.concatenate([k for r, tag, k, p in sorted(msgs, key=lambda m: m[0]) if tag == "in"])
    out = np.concatenate([k for r, tag, k, p in sorted(msgs, key=lambda m: m[0]) if tag == "out"])
    assert np.array_equal(out, oracle.scan(a, oracle.UINT, oracle.ULONG))


def test_rng_stream_partition():
    msgs = _run("rng", 3)
    parts = sorted((int(k[0]), int(k[1])) for r, tag, k, p in msgs)
    assert parts == [(0, 334), (334, 333), (667, 333)]
