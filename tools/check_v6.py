"""tools/check_v6.py -- development check of the keys-only onesweep kernel: sizes around tile
boundaries, u32/u64, against torch.sort; CLO_RADIX_PP_FLAGS=8 forces the repair path."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cl_ops_b200 as clo

ctx = clo.Context(); q = clo.Queue(ctx, stream=torch.cuda.current_stream().cuda_stream)
ok_all = True
for typ, tdt, nbits in ((clo.UINT, torch.int32, 32), (clo.ULONG, torch.int64, 64)):
    s = clo.CloSort("satradix", ctx, typ)
    for n in (1, 31, 8191, 8192, 8193, 100000, (1 << 20) + 77, (1 << 24) + 12345, 3 * 8192 * 300 + 5):
        g = torch.Generator(device="cuda"); g.manual_seed(n)
        lo, hi = (-2**31, 2**31 - 1) if nbits == 32 else (-2**63, 2**63 - 1)
        t_in = torch.randint(lo, hi, (n,), dtype=tdt, device="cuda", generator=g)
        if n > 50000:  # skew: many duplicates in the upper digits
            t_in[: n // 2] &= 0xFFFF
        t_out = torch.empty_like(t_in)
        bi, bo = clo.Buffer.wrap_tensor(ctx, t_in), clo.Buffer.wrap_tensor(ctx, t_out)
        s.with_device_data(q, bi, bo, n)
        torch.cuda.synchronize()
        # unsigned order reference
        if nbits == 32:
            ref = torch.sort(t_in.to(torch.int64) & 0xFFFFFFFF).values
            got = t_out.to(torch.int64) & 0xFFFFFFFF
        else:
            ref = torch.sort(t_in ^ (-2**63)).values ^ (-2**63)
            got = t_out
        ok = bool(torch.equal(ref, got))
        d = s.debug(q)
        print(nbits, n, "ok" if ok else "MISMATCH", "timeout", d[0], "repaired", d[1], flush=True)
        ok_all &= ok
        bi.destroy(); bo.destroy()
    s.destroy()
print("ALL OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
